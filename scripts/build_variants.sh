# geometry / option variants of the library for scripts/gpu_variants.sh: name:flags
set -e
cd "$(dirname "$0")/.."
rm -f xenomapper_b200/libxm_var_*.so
build() { # name flags...
  name=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -lz "$@" \
    -o xenomapper_b200/libxm_var_$name.so xenomapper_b200/csrc/xm_kernels.cu xenomapper_b200/csrc/xm_api.cu &
}
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  build $name $flags
done
wait
ls -la xenomapper_b200/libxm_var_*.so
