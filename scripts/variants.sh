# Tuning aid: builds variants of the native library with other compile-time settings into build/variants/ and
# times them on the GPU box against the default build (GMAPDP_LIB selects the library the engine loads).
#   usage (from the repo root, on a box with nvcc):  bash scripts/variants.sh "-DGMAPDP_GENOME_MINB=5" "-DGEN_TIECAP=1" ...
#   then on the B200:  GMAPDP_LIB=build/variants/lib_0.so python bench.py --no-cpu-baseline --chain-problems 0
mkdir -p build/variants
i=0
for flags in "$@"; do
  (cd gmap_2024_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $flags \
     -o ../../build/variants/lib_$i.so gmapdp_kernels.cu gmapdp_shim.cpp gmapdp_stream.cpp gmapchain_kernels.cu gmapchain_shim.cpp) && echo "build/variants/lib_$i.so: $flags"
  i=$((i+1))
done
