# Same-box A/B of the CTAs per SM the TMA LayerNorm kernels are compiled for: "TILE RING" pairs
for cfg in "$@"; do
  set -- $cfg
  export AVFE_EXTRA_NVCC="-DAVFE_FLT_TILE_CTAS=$1 -DAVFE_FLT_RING_CTAS=$2"
  echo "== tile kernel (16-bit) $1 CTAs/SM, ring kernel (float) $2 CTAs/SM"
  python -m avsl_b200.build --force > /dev/null 2>&1 || { echo "build failed"; continue; }
  python profiles/fuseln_sweep.py 2>&1 | tail -4
done
unset AVFE_EXTRA_NVCC
python -m avsl_b200.build --force > /dev/null 2>&1
