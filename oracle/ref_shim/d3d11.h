// d3d11.h (shim): empty on purpose, see stdafx.h
#pragma once
