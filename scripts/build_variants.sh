# Builds tuning variants of the library into scratch/variants/: bash scripts/build_variants.sh name "flags" [name "flags" ...]
set -e
rm -rf scratch/variants; mkdir -p scratch/variants
while [ $# -ge 2 ]; do
  LZB_SO=$PWD/scratch/variants/$1.so LZB_NVCC_EXTRA="$2" python lzfse_rust_b200/build.py --force > /dev/null
  echo "built $1 ($2)"
  shift 2
done
python lzfse_rust_b200/build.py --force > /dev/null
