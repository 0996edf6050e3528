# Build of the sm_100a shared library (C ABI) and the stand-alone bring-up tools.
# nvcc cross-compiles without a GPU; the built .so travels to the GPU box with the repo snapshot.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall
CSRC      := radvlm_b200/csrc
LIB       := radvlm_b200/libradvlm_b200.so
SRCS      := $(wildcard $(CSRC)/*.cu) $(wildcard $(CSRC)/*.cpp)
HDRS      := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/radvlm_b200.h
OBJS      := $(patsubst $(CSRC)/%,build/obj/%.o,$(SRCS))
TOOLS     := $(patsubst tools/%.cu,build/%,$(wildcard tools/*.cu))

all: $(LIB) tools

build/obj/%.cu.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build/obj
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

build/obj/%.cpp.o: $(CSRC)/%.cpp $(HDRS)
	@mkdir -p build/obj
	$(NVCC) $(NVCCFLAGS) -x cu -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS)

tools: $(TOOLS)

build/%: tools/%.cu $(LIB) include/radvlm_b200.h
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -o $@ $< -Lradvlm_b200 -lradvlm_b200 -Xlinker -rpath -Xlinker '$$ORIGIN/../radvlm_b200'

clean:
	rm -rf build $(LIB)

.PHONY: all tools clean
