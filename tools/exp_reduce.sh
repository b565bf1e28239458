for cfg in "2,2,2,2:4096" "3,3,3:4096" "3,3:8192" "3,4:4096" "3,2,2:4096" "2,3,3:4096" "3,3:16384" "4,4:4096" "2,4:8192" "3,3,2:2048"; do
  lr=${cfg%%:*}; d=${cfg##*:}
  echo "== LR=$lr DIRECT=$d"
  KZG_RED_LR=$lr KZG_RED_DIRECT=$d python tools/msm_phases.py 16 20 2>&1 | grep msm
done
