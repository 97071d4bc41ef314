# Builds libisdqn_b200.so (sm_100a only) in-tree.  `python -c "import __graft_entry__ as g; g.build()"` calls this.
NVCC ?= nvcc
PKG := is-dqn_b200
SRCS := $(wildcard $(PKG)/csrc/*.cu)
OBJS := $(patsubst $(PKG)/csrc/%.cu,build/%.o,$(SRCS))
HDRS := $(wildcard $(PKG)/csrc/*.cuh) include/isdqn_b200.h
NVCCFLAGS := $(EXTRA) -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
LIB := $(PKG)/lib/libisdqn_b200.so

all: $(LIB)

build/%.o: $(PKG)/csrc/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	@mkdir -p $(PKG)/lib
	$(NVCC) -shared -o $@ $(OBJS) -ldl -Xlinker --version-script=$(PKG)/csrc/exports.map

clean:
	rm -rf build $(LIB)
.PHONY: all clean
