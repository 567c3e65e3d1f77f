# Builds tuning variants of the library into scratch/variants/.
set -e
rm -rf scratch/variants; mkdir -p scratch/variants
build() {  # name, flags
  LZB_SO=$PWD/scratch/variants/$1.so LZB_NVCC_EXTRA="$2" python lzfse_rust_b200/build.py --force | grep -E "k_enc_parse" -A2 | grep -E "Used|spill" | tr '\n' ' '
  echo " <- $1"
}
build sink_smem_c7 "-DLZB_PARSE_SINK_SMEM=1 -DLZB_PARSE_CTAS=7"
build sink_smem_c8 "-DLZB_PARSE_SINK_SMEM=1 -DLZB_PARSE_CTAS=8"
build sink_stack_c7 "-DLZB_PARSE_SINK_SMEM=0 -DLZB_PARSE_CTAS=7"
python lzfse_rust_b200/build.py --force > /dev/null
