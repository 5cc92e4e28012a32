export BVG_DEBUG_BUILD=1
timeout 300 python -m pytest tests/test_gpu_act_tc.py -x -q -m gpu 2>&1 | tail -3
for d in 0 15 3 12; do
  echo "== dry=$d"; BVG_ACT_TC_DRY=$d timeout 100 python tools/act_tc_roles.py 96,60160,32 2>&1 | tail -1
done
BVG_ACT_TC_DRY=0 timeout 100 python tools/act_tc_roles.py 2>&1 | tail -6
