#pragma once
namespace cooperative_groups {}
