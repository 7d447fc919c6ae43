// stdafx.h (shim) -- TEST INFRASTRUCTURE ONLY.  Stands in for the reference's precompiled header
// (src/pg/pg1_embree/stdafx.h: Win32 / D3D11 / Dear ImGui / Embree includes) so that the reference's OWN, UNMODIFIED
// pg1/*.cpp files compile with g++ where they lie under /root/reference (oracle/ref.mk builds oracle/_ref/libpg_ref.so).
// Nothing of the reference is copied: this header only supplies what the platform headers would have declared.
#pragma once
#include <stdio.h>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cfloat>
#include <cmath>
// MSVC declares the float overloads of the <cmath> functions in the GLOBAL namespace (corecrt_math.h, C++ mode), so the
// reference's unqualified sqrt(float) / atan2(float, float) / asin(float) / floor(float) / abs(float) calls in files without
// `using namespace std` (vector3.cpp:16,33; SphericalMap.cpp:22-23; texture.cpp:88-91) pick them.  With libstdc++ only
// <math.h> (not <cmath>) brings those overloads into the global namespace; without this line g++ would evaluate them in double.
#include <math.h>
#include <stdlib.h>
#include <string>
#include <chrono>
#include <mutex>
#include <thread>
#include <atomic>
#include <vector>
#include <map>
#include <memory>
#include <random>
#include <functional>
#include <stdexcept>
#include <iostream>
#include <algorithm>
#include <cassert>
#include <sys/types.h>
#include <xmmintrin.h>
#include <pmmintrin.h>

// Intel Embree 3 API: the reference's vendored headers (declarations only; oracle/ref_shim/ref_embree.cpp defines them)
#include <embree3/rtcore.h>

// ---- Win32 vocabulary the class declarations mention (simpleguidx11.h); no Win32 function is ever called
typedef void* HWND; typedef void* HINSTANCE; typedef long HRESULT; typedef long LRESULT; typedef unsigned int UINT;
typedef uintptr_t WPARAM; typedef intptr_t LPARAM;
#define CALLBACK
#define S_OK 0L
struct WNDCLASSEX { int unused; };
struct ID3D11Device; struct ID3D11DeviceContext; struct IDXGISwapChain; struct ID3D11RenderTargetView;
struct ID3D11Texture2D; struct ID3D11ShaderResourceView;

// ---- Dear ImGui: the handful of calls Raytracer::Ui makes (raytracer.cpp:448-507), as no-ops
struct ImGuiIO { float Framerate = 1.0f; };
namespace ImGui {
inline bool Begin(const char*, bool* = nullptr, int = 0) { return true; }
inline void End() {}
inline void Text(const char*, ...) {}
inline void Separator() {}
inline bool Checkbox(const char*, bool*) { return false; }
inline bool SliderFloat(const char*, float*, float, float) { return false; }   // leaves the slider value (0.5, raytracer.cpp:450) untouched
inline ImGuiIO& GetIO() { static ImGuiIO io; return io; }
}

// ---- MSVC-isms in the sources
#define _CMATH_ std                       /* PinHoleCamera.cpp:16  _CMATH_::tan      */
#define _fseeki64 fseeko                  /* utils.cpp:38-40                            */
#define _ftelli64 ftello
#define sscanf_s sscanf
#define _USE_MATH_DEFINES
// windows.h's min / max macros (vector3.cpp:53 relies on them); defined last, after every standard header above
#ifndef max
#define max(a, b) (((a) > (b)) ? (a) : (b))
#endif
#ifndef min
#define min(a, b) (((a) < (b)) ? (a) : (b))
#endif
