# Builds the C-ABI library (include/dzahui_b200.h) for sm_100a and the CPU oracle (test infrastructure).
# Same flags as dzahui_b200/build.py (which the Python side, the tests and __graft_entry__.build() use).
NVCC ?= /usr/local/cuda/bin/nvcc
HOSTCXX ?= $(if $(wildcard /usr/bin/g++),/usr/bin/g++,g++)
CSRC := dzahui_b200/csrc
LIB := dzahui_b200/libdzahui_b200.so
SRCS := $(CSRC)/dzb_api.cu $(CSRC)/gl_quadrature.cpp $(CSRC)/mesh_ingest.cpp $(CSRC)/csv_writer.cpp
HDRS := $(wildcard $(CSRC)/*.cuh $(CSRC)/*.h) include/dzahui_b200.h

all: $(LIB) oracle

$(LIB): $(SRCS) $(HDRS)
	$(NVCC) -ccbin $(HOSTCXX) -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
	    -Xcompiler -fPIC,-ffp-contract=off,-O2 -shared -o $@ $(SRCS)

oracle:
	$(MAKE) -s -C oracle

host_mirror: $(LIB) tests/cpp/host_mirror.cpp include/dzahui_b200.hpp
	$(HOSTCXX) -std=c++17 -O1 -Wall -I include tests/cpp/host_mirror.cpp -o tests/cpp/host_mirror \
	    -L dzahui_b200 -ldzahui_b200 -Wl,-rpath,$(abspath dzahui_b200)

clean:
	rm -f $(LIB) tests/cpp/host_mirror
	$(MAKE) -s -C oracle clean

.PHONY: all oracle clean host_mirror
