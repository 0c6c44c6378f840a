// TEST INFRASTRUCTURE: stand-in for the un-vendored header of this name (see scene_shim.h)
#pragma once
#include "../scene_shim.h"
