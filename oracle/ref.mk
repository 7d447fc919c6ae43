# oracle/ref.mk -- builds oracle/_ref/libpg_ref.so: the reference's OWN pg1/*.cpp files, compiled UNMODIFIED from where they lie
# under /root/reference, plus the stubs of oracle/ref_shim/ for what the reference does not vendor (Embree and FreeImage binaries,
# the Win32 / D3D11 / ImGui window).  Test infrastructure only; outputs only under oracle/_ref/ (git-ignored, shipped to the GPU box).
#   make -f ref.mk            (run from oracle/; a no-op with a message when /root/reference is absent)
# No reference source is copied: oracle/_ref/src/ holds SYMLINKS to the reference's files next to the shim headers, because a
# quoted #include looks in the including file's own directory first and the reference's stdafx.h (UTF-16, Windows-only) must
# not be the one that is found.
REF ?= /root/reference/src/pg/pg1_embree
LIBS ?= /root/reference/src/libs
ORC_CXX := $(shell test -x /usr/bin/g++ && echo /usr/bin/g++ || echo g++)
# -ffp-contract=off: MSVC /fp:precise x64 does not contract; -fpermissive -w: MSVC dialect (implicit narrowing, %I64u, ...)
REF_CXXFLAGS = -O2 -std=c++17 -ffp-contract=off -fno-fast-math -fopenmp -fPIC -fpermissive -w -I $(LIBS)/embree/include -I $(LIBS)/freeimage/include
REF_UNITS = raytracer PinHoleCamera LightSource SphericalMap texture utils vector3 matrix3x3 material surface triangle vertex structs mymath objloader tutorials
SHIM_UNITS = ref_embree ref_freeimage ref_gui ref_api

ifeq ($(wildcard $(REF)/raytracer.cpp),)
all:
	@echo "oracle/ref.mk: $(REF) is absent (GPU box): keeping the prebuilt oracle/_ref/libpg_ref.so"
else
all: _ref/libpg_ref.so

_ref/src/.stamp: $(wildcard ref_shim/*.h)
	rm -rf _ref/src && mkdir -p _ref/src _ref/obj
	for f in $(REF)/*.cpp $(REF)/*.h; do b=$$(basename $$f); [ "$$b" = stdafx.h ] || [ "$$b" = stdafx.cpp ] || ln -s $$f _ref/src/$$b; done
	for f in ref_shim/*.h ref_shim/*.cpp; do ln -sf ../../$$f _ref/src/$$(basename $$f); done
	touch $@

_ref/obj/%.o: _ref/src/.stamp
	$(ORC_CXX) $(REF_CXXFLAGS) -c _ref/src/$*.cpp -o $@

_ref/libpg_ref.so: $(addprefix _ref/obj/,$(addsuffix .o,$(REF_UNITS) $(SHIM_UNITS))) libpg_oracle.so
	$(ORC_CXX) -shared -fopenmp -o $@ $(filter %.o,$^) -L. -lpg_oracle -Wl,-rpath,'$$ORIGIN/..'
endif

clean:
	rm -rf _ref
