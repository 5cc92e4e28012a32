# geometry sweep of the input staging of act1d_tc_kernel (rebuilds the library on the GPU box for each configuration)
export BVG_DEBUG_BUILD=1
for cfg in "-DBVG_TXBLK=4 -DBVG_TXSTAGES=3" "-DBVG_TXBLK=2 -DBVG_TXSTAGES=6" "-DBVG_TXBLK=1 -DBVG_TXSTAGES=10" "-DBVG_TXBLK=2 -DBVG_TXSTAGES=3"; do
  export BVG_EXTRA_NVCC="$cfg"
  echo "==== $cfg"
  for d in 15 0; do
    echo "-- dry=$d"; BVG_ACT_TC_DRY=$d timeout 300 python tools/act_tc_roles.py 96,60160,32 192,15040,32 2>&1 | tail -2
  done
done
