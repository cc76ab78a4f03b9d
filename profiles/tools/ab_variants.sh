# A/B of experiment builds of libdvsloss (deep-visual-slam_b200/csrc/build.py --variant NAME -D...), one bench line each.
for v in "$@"; do
  f=deep-visual-slam_b200/dvsloss/libdvsloss$v.so
  DVSLOSS_LIB=$PWD/$f python bench.py --no-cpu --no-eager --no-train 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('VARIANT', '$v', 'ms_per_step', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4))"
done
