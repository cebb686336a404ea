// TEST INFRASTRUCTURE: stand-in for the un-vendored header of this name (see ../../ls2d_ref_shim.h)
#pragma once
#include "../../ls2d_ref_shim.h"
