# Same-box A/B of lip_frame_kernel's warp split / ring depth / footprint slots: rebuilds the library per
# configuration (-D knobs of avfe_lip_frame.cuh) and times the lip stage (profiles/lip_shapes.py sweep).
#   bash profiles/lip_sweep.sh "P88 R88 S88 P96 R96 S96" ...     (one quoted sextuple per configuration)
for cfg in "$@"; do
  set -- $cfg
  export AVFE_EXTRA_NVCC="-DAVFE_LIP_PHASES_88=$1 -DAVFE_LIP_RING_88=$2 -DAVFE_LIP_SLOTS_88=$3 -DAVFE_LIP_PHASES_96=$4 -DAVFE_LIP_RING_96=$5 -DAVFE_LIP_SLOTS_96=$6 $AVFE_SWEEP_MORE"
  echo "== phases/ring/slots 88: $1/$2/$3   96: $4/$5/$6"
  python -m avsl_b200.build --force > /dev/null 2>&1 || { echo "build failed"; continue; }
  python profiles/lip_shapes.py sweep 2>&1 | tail -3
done
unset AVFE_EXTRA_NVCC
python -m avsl_b200.build --force > /dev/null 2>&1
