// ref_gui.cpp -- TEST INFRASTRUCTURE ONLY.  Members of SimpleGuiDX11 (pg1/simpleguidx11.h) that the reference defines in
// simpleguidx11.cpp next to its Win32 / D3D11 / ImGui window code, which cannot be compiled here.  Only what Raytracer's
// vtable and constructor need: no window, no producer thread (oracle/ref_shim/ref_api.cpp drives the pixel loop itself).
#include "stdafx.h"
#include "simpleguidx11.h"

SimpleGuiDX11::SimpleGuiDX11(const int width, const int height) { width_ = width; height_ = height; }   // simpleguidx11.cpp:4-10 without Init()
SimpleGuiDX11::~SimpleGuiDX11() {}
int SimpleGuiDX11::MainLoop() { return 0; }
int SimpleGuiDX11::Ui() { return 0; }                                                                      // simpleguidx11.cpp:76-79
Color4f SimpleGuiDX11::get_pixel(const int, const int, const float) { return Color4f{1.0f, 0.0f, 1.0f, 1.0f}; }   // :81-84
int SimpleGuiDX11::width() const { return width_; }
int SimpleGuiDX11::height() const { return height_; }
