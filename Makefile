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
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -shared -o $@ $(OBJS) -ldl -Xlinker --version-script=$(PKG)/csrc/exports.map

# optional: the XLA FFI handlers (needs xla/ffi/api/ffi.h, e.g. XLA_FFI_INCLUDE=$$(python -c "import jaxlib,os;print(os.path.join(jaxlib.__path__[0],'include'))"))
ffi: $(LIB)
	g++ -O2 -std=c++17 -fPIC -shared -I$(XLA_FFI_INCLUDE) -I/usr/local/cuda/include $(PKG)/csrc/xla_ffi.cc -o $(PKG)/lib/libisdqn_b200_ffi.so -L$(PKG)/lib -lisdqn_b200

clean:
	rm -rf build $(LIB)
.PHONY: all clean ffi
