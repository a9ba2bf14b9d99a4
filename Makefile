# Build libhowl_b200.so (sm_100a) in-tree, plus the oracle's nothing-to-compile marker.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall
SRC := $(wildcard howl_b200/csrc/*.cu)
OBJ := $(patsubst howl_b200/csrc/%.cu,build/%.o,$(SRC))
LIB := howl_b200/lib/libhowl_b200.so

all: $(LIB)

build/%.o: howl_b200/csrc/%.cu $(wildcard howl_b200/csrc/*.cuh) include/howl_b200.h include/howl_b200_debug.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	@mkdir -p howl_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart

clean:
	rm -rf build $(LIB)
.PHONY: all clean
