# Builds libqb2.so (the CUDA engine behind include/qrisp_b200.h) for sm_100a.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC -Iinclude -Iqrisp_b200/csrc $(EXTRA)
CSRC      := qrisp_b200/csrc
SRC       := $(CSRC)/qb2_api.cu $(CSRC)/qb2_pass.cu $(CSRC)/qb2_measure.cu $(CSRC)/qb2_prefix.cu $(CSRC)/qb2_block_tc.cu
# one translation unit per pass-kernel variant (qb2_pass_inst.cu with -DQB2_INST=k): they build in parallel
INST      := 0 1 2 3 4 5 6 7 8 9 10 11 12 13
INSTOBJ   := $(foreach k,$(INST),$(CSRC)/qb2_pass_inst_$(k).o)
OBJ       := $(SRC:.cu=.o) $(INSTOBJ)
HDR       := include/qrisp_b200.h $(CSRC)/qb2_internal.cuh
PASSHDR   := $(CSRC)/qb2_pass_common.cuh $(CSRC)/qb2_pass_kernel.cuh $(CSRC)/qb2_pass_tma.cuh $(CSRC)/qb2_pass_ops.inc
LIB       := qrisp_b200/lib/libqb2.so

all: $(LIB)

$(CSRC)/qb2_pass_inst_%.o: $(CSRC)/qb2_pass_inst.cu $(HDR) $(PASSHDR)
	$(NVCC) $(NVCCFLAGS) $(PTXAS) -DQB2_INST=$* -c $< -o $@

$(CSRC)/%.o: $(CSRC)/%.cu $(HDR)
	$(NVCC) $(NVCCFLAGS) $(PTXAS) -c $< -o $@

$(LIB): $(OBJ)
	mkdir -p qrisp_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ)

clean:
	rm -f $(OBJ) $(LIB)

.PHONY: all clean
