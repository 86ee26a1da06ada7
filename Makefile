# Builds libqb2.so (the CUDA engine behind include/qrisp_b200.h) for sm_100a, and the C oracle.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC -Iinclude -Iqrisp_b200/csrc $(EXTRA)
SRC       := qrisp_b200/csrc/qb2_api.cu qrisp_b200/csrc/qb2_pass.cu qrisp_b200/csrc/qb2_measure.cu
OBJ       := $(SRC:.cu=.o)
LIB       := qrisp_b200/lib/libqb2.so

all: $(LIB)

%.o: %.cu include/qrisp_b200.h qrisp_b200/csrc/qb2_internal.cuh
	$(NVCC) $(NVCCFLAGS) $(PTXAS) -c $< -o $@

$(LIB): $(OBJ)
	mkdir -p qrisp_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ)

# development build: one kernel variant (complex64, r=5) with time stamps along a tile's life
# (tools/trace_tile.py)
trace:
	mkdir -p qrisp_b200/lib
	$(NVCC) $(NVCCFLAGS) -DQB2_TRACE -DQB2_ONLY_F5 -shared -o qrisp_b200/lib/libqb2_trace.so $(SRC)

clean:
	rm -f $(OBJ) $(LIB)

.PHONY: all clean trace
