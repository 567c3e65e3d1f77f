# Builds tuning variants of the library into scratch/variants/ (expand kernel: ring size, workers, CTAs per SM, wait back-off).
set -e
rm -rf scratch/variants; mkdir -p scratch/variants
build() {  # name, flags
  LZB_SO=$PWD/scratch/variants/$1.so LZB_NVCC_EXTRA="$2" python lzfse_rust_b200/build.py --force | grep -E "k_expand_cta" -A2 | grep -E "Used|spill" | tr '\n' ' '
  echo " <- $1"
}
build r16w7c3 "-DLZB_XWORKERS=7"
build r16w7c3s20 "-DLZB_XWORKERS=7 -DLZB_XSLEEP=20"
build r16w5c3 "-DLZB_XWORKERS=5"
build r16w3c3 "-DLZB_XWORKERS=3"
build r15w3c6 "-DLZB_XRING_LOG2=15 -DLZB_XWORKERS=3 -DLZB_XCTAS=6"
build r15w5c4 "-DLZB_XRING_LOG2=15 -DLZB_XWORKERS=5 -DLZB_XCTAS=4"
python lzfse_rust_b200/build.py --force > /dev/null
