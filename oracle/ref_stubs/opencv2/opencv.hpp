// stub: RVI/parameter/parameters.h includes OpenCV for types the factor sources never use (and, implicitly, for the
// standard headers below)
#pragma once
#include <string>
#include <vector>
