#pragma once
/* stand-in for nperf.h (common.h:6) */
