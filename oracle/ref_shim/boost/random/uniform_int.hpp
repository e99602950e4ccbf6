// TEST INFRASTRUCTURE ONLY (oracle/): stand-in, see ../random.hpp
#pragma once
#include "../random.hpp"
