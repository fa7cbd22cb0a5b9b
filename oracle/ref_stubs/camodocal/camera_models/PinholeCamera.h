// TEST INFRASTRUCTURE stub
#pragma once
#include "CameraFactory.h"
