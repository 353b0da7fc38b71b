#pragma once
/* stand-in for the LSU "gp" library header (not in the reference tree, common.h:4) */
#include <map>
#include <set>
#include <string>
