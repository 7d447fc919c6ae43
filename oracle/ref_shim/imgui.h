// imgui.h (shim): empty on purpose, see stdafx.h
#pragma once
