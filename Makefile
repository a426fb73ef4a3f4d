# Builds the product: libvstree_b200.so (CUDA kernels + C ABI, sm_100a only), the mkvtree host driver and the mkcld tool.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v
CSRC := vstree_b200/csrc
LIBDIR := vstree_b200/lib
BINDIR := vstree_b200/bin
OBJDIR := build/obj
CU := $(wildcard $(CSRC)/*.cu)
OBJ := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(CU))
HDR := $(wildcard $(CSRC)/*.cuh) include/vstree_b200.h

all: $(LIBDIR)/libvstree_b200.so $(BINDIR)/mkvtree $(BINDIR)/mkcld

# the drop-in host driver: plain C, links the C ABI only
$(BINDIR)/mkvtree: vstree_b200/host/mkvtree.c include/vstree_b200.h $(LIBDIR)/libvstree_b200.so
	@mkdir -p $(BINDIR)
	gcc -O2 -std=gnu99 -Wall -Wextra -Iinclude -o $@ vstree_b200/host/mkvtree.c -L$(LIBDIR) -lvstree_b200 -lz -lpthread \
	    -Wl,-rpath,'$$ORIGIN/../lib'

# the child-table tool (reference: src/Mkvtree/mkcld.mn.c): plain C on the C ABI
$(BINDIR)/mkcld: vstree_b200/host/mkcld.c include/vstree_b200.h $(LIBDIR)/libvstree_b200.so
	@mkdir -p $(BINDIR)
	gcc -O2 -std=gnu99 -Wall -Wextra -Iinclude -o $@ vstree_b200/host/mkcld.c -L$(LIBDIR) -lvstree_b200 \
	    -Wl,-rpath,'$$ORIGIN/../lib'

$(OBJDIR)/%.o: $(CSRC)/%.cu $(HDR)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; exit 1)

$(LIBDIR)/libvstree_b200.so: $(OBJ)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ)

clean:
	rm -rf build $(LIBDIR) $(BINDIR)

.PHONY: all clean
