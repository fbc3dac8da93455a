# Top-level build: CUDA library (sm_100a only), host C++ mirror + CLI, oracle (test infra).
#   make            -> cobs_b200/lib/libcobsgpu.so, build/ host binaries, oracle/liboracle.so
#   make ref        -> oracle/_ref/ (needs /root/reference)
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       := g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function,-Wno-unknown-pragmas \
             -Xptxas -v --expt-relaxed-constexpr
CSRC      := cobs_b200/csrc
LIB       := cobs_b200/lib/libcobsgpu.so

all: $(LIB) host oracle build/kernel_unit_tests

$(LIB): $(CSRC)/cobsgpu.cu $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.hpp) include/cobsgpu.h
	@mkdir -p cobs_b200/lib build
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)/cobsgpu.cu 2> build/ptxas.log || (cat build/ptxas.log; false)
	@grep -E "error|warning" build/ptxas.log | grep -v "ptxas info" || true

HOST_SRCS := $(wildcard cobs_b200/host/src/*.cpp)
HOST_INC  := -Icobs_b200/host/include -Iinclude
HOST_FLAGS := -std=c++17 -O2 -Wall -fPIC $(HOST_INC)

ifneq ($(wildcard cobs_b200/host/src/*.cpp),)
host: build/libcobs_b200.so build/cobs build/host_tests build/host_unit_tests
else
host:
	@echo 'host sources not present yet'
endif

build/libcobs_b200.so: $(HOST_SRCS) $(LIB) $(shell find cobs_b200/host/include -name '*.hpp' 2>/dev/null)
	@mkdir -p build
	$(CXX) $(HOST_FLAGS) -shared -o $@ $(HOST_SRCS) -Lcobs_b200/lib -lcobsgpu \
	    -Wl,-rpath,'$$ORIGIN/../cobs_b200/lib' -lpthread

build/cobs: cobs_b200/host/cli/cobs_main.cpp build/libcobs_b200.so
	$(CXX) $(HOST_FLAGS) -o $@ $< -Lbuild -lcobs_b200 -Lcobs_b200/lib -lcobsgpu \
	    -Wl,-rpath,'$$ORIGIN' -Wl,-rpath,'$$ORIGIN/../cobs_b200/lib'

build/host_tests: cobs_b200/host/tests/host_tests.cpp build/libcobs_b200.so
	$(CXX) $(HOST_FLAGS) -o $@ $< -Lbuild -lcobs_b200 -Lcobs_b200/lib -lcobsgpu \
	    -Wl,-rpath,'$$ORIGIN' -Wl,-rpath,'$$ORIGIN/../cobs_b200/lib'

build/host_unit_tests: cobs_b200/host/tests/host_unit_tests.cpp build/libcobs_b200.so
	$(CXX) $(HOST_FLAGS) -o $@ $< -Lbuild -lcobs_b200 -Lcobs_b200/lib -lcobsgpu \
	    -Wl,-rpath,'$$ORIGIN' -Wl,-rpath,'$$ORIGIN/../cobs_b200/lib'

# host-side unit tests of the __host__ __device__ kernel arithmetic (test infrastructure: links
# the oracle)
build/kernel_unit_tests: tests/csrc/kernel_unit_tests.cu $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.hpp) oracle
	@mkdir -p build
	$(NVCC) $(ARCH) -O2 -std=c++17 --expt-relaxed-constexpr --extended-lambda -Xcompiler -Wno-unknown-pragmas \
	    -o $@ $< -Loracle -loracle -Xlinker -rpath -Xlinker '$$ORIGIN/../oracle' 2> build/kut.log || (cat build/kut.log; false)

oracle:
	$(MAKE) -s -C oracle

ref:
	$(MAKE) -s -C oracle -j8 ref

clean:
	rm -rf build cobs_b200/lib/*.so
	$(MAKE) -C oracle clean

.PHONY: all host oracle ref clean
