# Builds the product: ntcard_b200/libntcard_b200.so (C-ABI + sm_100a kernels) and the host CLIs bin/ntcard, bin/nthll.
# The checkers under oracle/ have their own Makefile.
NVCC ?= /usr/local/cuda/bin/nvcc
CXX = g++
ARCH = -gencode arch=compute_100a,code=sm_100a
NVFLAGS = -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-Wextra,-Wno-unused-parameter -Xptxas -v --expt-relaxed-constexpr
CXXFLAGS = -O3 -std=c++17 -fPIC -Wall -Wextra -ffp-contract=off
SRC = ntcard_b200/csrc
BUILD = build

# k mod 31 variants of the bit-sliced scan kernel to compile (each is one translation unit, ~3 s and 0.45 MB): all of them, so that
# every k < 288 takes the pipeline
BS_KMS ?= 0 1 2 3 4 5 6 7 8 9 10 11 12 13 14 15 16 17 18 19 20 21 22 23 24 25 26 27 28 29 30
BS_KM_LIST = $(foreach n,$(BS_KMS),X($(n)))
CU_SRCS = $(SRC)/ntc_api.cu $(SRC)/sketch_kernels.cu $(SRC)/hit_kernels.cu $(SRC)/hll_kernels.cu $(SRC)/bitslice_dispatch.cu
BS_OBJS = $(foreach n,$(BS_KMS),$(BUILD)/bitslice_km$(n).o)
CPP_SRCS = $(SRC)/host_util.cpp
CU_OBJS = $(patsubst $(SRC)/%.cu,$(BUILD)/%.o,$(CU_SRCS))
CPP_OBJS = $(patsubst $(SRC)/%.cpp,$(BUILD)/%.o,$(CPP_SRCS))
HDRS = $(wildcard $(SRC)/*.h $(SRC)/*.cuh include/*.h)
LIB = ntcard_b200/libntcard_b200.so

all: $(LIB) cli

$(BUILD)/bitslice_km%.o: $(SRC)/bitslice_inst.cu $(HDRS)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -DBS_KM=$* -c $< -o $@ 2> $(BUILD)/bitslice_km$*.ptxas.log || (cat $(BUILD)/bitslice_km$*.ptxas.log; exit 1)
	@grep -E "error|warning|spill|registers" $(BUILD)/bitslice_km$*.ptxas.log | grep -v "0 bytes spill" | head -8 || true

$(BUILD)/bitslice_dispatch.o: $(SRC)/bitslice_dispatch.cu $(HDRS) Makefile
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) '-DBS_KM_LIST=$(BS_KM_LIST)' -c $< -o $@

$(BUILD)/%.o: $(SRC)/%.cu $(HDRS)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(BUILD)/$*.ptxas.log || (cat $(BUILD)/$*.ptxas.log; exit 1)
	@grep -E "error|warning|spill|registers" $(BUILD)/$*.ptxas.log | grep -v "0 bytes spill" | head -40 || true

$(BUILD)/%.o: $(SRC)/%.cpp $(HDRS)
	@mkdir -p $(BUILD)
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(LIB): $(CU_OBJS) $(BS_OBJS) $(CPP_OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $^ -cudart static -Xlinker --exclude-libs,ALL

CLI_LINK = -Lntcard_b200 -lntcard_b200 -Wl,-rpath,'$$ORIGIN/../ntcard_b200' -lpthread
cli: bin/ntcard bin/nthll
bin/ntcard: $(SRC)/cli/ntcard_main.cpp $(SRC)/cli/reader.cpp $(SRC)/cli/reader.h $(LIB) $(HDRS)
	@mkdir -p bin
	$(CXX) $(CXXFLAGS) -fopenmp -Iinclude -o $@ $(SRC)/cli/ntcard_main.cpp $(SRC)/cli/reader.cpp $(CLI_LINK)
bin/nthll: $(SRC)/cli/nthll_main.cpp $(SRC)/cli/reader.cpp $(SRC)/cli/reader.h $(LIB) $(HDRS)
	@mkdir -p bin
	$(CXX) $(CXXFLAGS) -Iinclude -o $@ $(SRC)/cli/nthll_main.cpp $(SRC)/cli/reader.cpp $(CLI_LINK)

clean:
	rm -rf $(BUILD) $(LIB) bin

.PHONY: all cli clean
