// freeimage.h (shim): texture.h:4 includes the header by its lower-case name (Windows file systems do not care);
// the declarations are FreeImage's own, vendored by the reference; oracle/ref_shim/ref_freeimage.cpp defines the calls texture.cpp makes
#pragma once
#include <FreeImage.h>
