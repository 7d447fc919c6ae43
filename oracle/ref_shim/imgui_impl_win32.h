// imgui_impl_win32.h (shim): empty on purpose, see stdafx.h
#pragma once
