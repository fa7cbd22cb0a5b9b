// TEST INFRASTRUCTURE stub (see opencv.hpp)
#pragma once
#include "opencv.hpp"
