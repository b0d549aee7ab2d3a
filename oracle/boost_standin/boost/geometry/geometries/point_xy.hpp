// Stand-in header: see ../kcbs_standin.hpp (test infrastructure of the oracle).
#pragma once
#include <boost/geometry/kcbs_standin.hpp>
