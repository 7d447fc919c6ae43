// imgui_impl_dx11.h (shim): empty on purpose, see stdafx.h
#pragma once
