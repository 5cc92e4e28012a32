# runtime-switch sweep of actconv_tc_kernel (ring depths / conv accumulator stages)
for cfg in "2 0 0" "1 0 0"; do
  set -- $cfg
  echo "==== NACC=$1 NU=$2 NY=$3"
  BVG_TCF_NACC=$1 BVG_TCF_NU=$2 BVG_TCF_NY=$3 timeout 200 python tools/actconv_tc_roles.py quick 2>&1 | cut -c1-330
done
