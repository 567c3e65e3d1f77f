# Builds tuning variants of the library into scratch/variants/.
set -e
rm -rf scratch/variants; mkdir -p scratch/variants
build() {  # name, flags
  LZB_SO=$PWD/scratch/variants/$1.so LZB_NVCC_EXTRA="$2" python lzfse_rust_b200/build.py --force | grep -E "k_expandE" -A2 | grep -E "Used|spill" | tr '\n' ' '
  echo " <- $1"
}
build exp_c4 "-DLZB_EXPAND_CTAS=4"
build exp_c5 "-DLZB_EXPAND_CTAS=5"
build exp_c6 "-DLZB_EXPAND_CTAS=6"
python lzfse_rust_b200/build.py --force > /dev/null
