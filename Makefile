# Makefile -- the native build for C/C++ users (python -m csgn_b200.build does the same from Python).
#
#   make                     libcsgn.so (sm_100a kernels + C ABI) and libcertFHE.so (the certFHE classes over it)
#   make testers REFERENCE=/path/to/certFHE   the reference's own tests/*.cpp, UNMODIFIED, against this library
#                            (tester_basic_operations, tester_permutations, tester_timings -- the targets of the
#                            reference's CMakeLists.txt:45-52); nothing is copied, the sources are compiled in place
#   make cpp-tests           tests/cpp/accept_demos and sharded_demo
#   make tools               tools/bin/cpp_e2e (the bench step through the C++ API; bench.py reports it as e2e_cpp)
#   make variants            libcsgn_variants.so: the losing kernel variants of the tuning sweeps compiled in
#   make clean
#
# nvcc cross-compiles for sm_100a without a GPU.  There is no CPU build of the kernels: no B200, no library.
NVCC      ?= nvcc
CXX       ?= g++
LIBDIR    := csgn_b200/lib
CSRC      := $(wildcard csgn_b200/csrc/*.cu)
CHDR      := $(wildcard csgn_b200/csrc/*.cuh) include/csgn.h
FHESRC    := $(wildcard csgn_b200/certfhe/*.cpp)
FHEHDR    := $(wildcard csgn_b200/certfhe/*.h)
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC
REFERENCE ?= /root/reference
DROPIN    := build/dropin

all: $(LIBDIR)/libcsgn.so $(LIBDIR)/libcertFHE.so

$(LIBDIR)/libcsgn.so: $(CSRC) $(CHDR)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(CSRC)

$(LIBDIR)/libcertFHE.so: $(FHESRC) $(FHEHDR) include/csgn.h $(LIBDIR)/libcsgn.so
	$(CXX) -O3 -std=c++11 -fPIC -shared -Wall -Iinclude -Icsgn_b200/certfhe -o $@ $(FHESRC) -L$(LIBDIR) -lcsgn '-Wl,-rpath,$$ORIGIN'

# the reference's demo programs say #include "../src/certFHE.h": a symlink tree makes that land on our header
testers: all
	@mkdir -p $(DROPIN)/tests $(DROPIN)/bin
	@[ -L $(DROPIN)/src ] || ln -s $(abspath csgn_b200/certfhe) $(DROPIN)/src
	@for f in $(REFERENCE)/tests/*.cpp; do n=$$(basename $$f .cpp); \
	  [ -L $(DROPIN)/tests/$$n.cpp ] || ln -s $$f $(DROPIN)/tests/$$n.cpp; \
	  echo "$(CXX) tester_$$n"; \
	  $(CXX) -O2 -std=c++11 -w -Iinclude -o $(DROPIN)/bin/tester_$$n $(DROPIN)/tests/$$n.cpp -L$(LIBDIR) -lcertFHE -lcsgn \
	    '-Wl,-rpath,$$ORIGIN/../../../csgn_b200/lib' || exit 1; done

cpp-tests: all
	@mkdir -p tests/cpp/bin
	@for n in accept_demos sharded_demo; do echo "$(CXX) $$n"; \
	  $(CXX) -O2 -std=c++11 -Wall -Icsgn_b200/certfhe -Iinclude -o tests/cpp/bin/$$n tests/cpp/$$n.cpp -L$(LIBDIR) -lcertFHE -lcsgn \
	    '-Wl,-rpath,$$ORIGIN/../../../csgn_b200/lib' || exit 1; done

tools: all
	@mkdir -p tools/bin
	for n in cpp_e2e ctor_probe; do \
	  $(CXX) -O2 -std=c++11 -Wall -Icsgn_b200/certfhe -Iinclude -o tools/bin/$$n tools/$$n.cpp -L$(LIBDIR) -lcertFHE -lcsgn \
	      '-Wl,-rpath,$$ORIGIN/../../csgn_b200/lib' || exit 1; \
	done

variants: $(CSRC) $(CHDR)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVCCFLAGS) -DCSGN_BUILD_VARIANTS -shared -o $(LIBDIR)/libcsgn_variants.so $(CSRC)

clean:
	rm -rf $(LIBDIR)/*.so build tests/cpp/bin tools/bin

.PHONY: all testers cpp-tests tools variants clean
