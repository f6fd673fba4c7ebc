// TEST INFRASTRUCTURE ONLY: see ../random.hpp
#include "../random.hpp"
