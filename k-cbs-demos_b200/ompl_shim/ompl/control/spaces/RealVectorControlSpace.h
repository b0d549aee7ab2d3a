// Forwarding header of the OMPL API shim (see kcbs_shim.h).
#pragma once
#include "../../kcbs_shim.h"
