// corecrt_math_defines.h (shim): the MSVC header SphericalMap.cpp:4 includes for M_PI
#pragma once
#include <cmath>
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
