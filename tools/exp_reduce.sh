for w in 32 64 128; do for v in 0 1 2; do
  echo "== TAIL_WIDTH=$w L0_VARIANT=$v"
  KZGB200_TAIL_WIDTH=$w KZGB200_L0_VARIANT=$v python tools/msm_phases.py 16 20 2>&1 | grep msm
done; done
