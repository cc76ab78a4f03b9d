# DDP option A/B of the training step at N GPUs: bash profiles/tools/ddp_opts.sh 2 "" "broadcast_buffers=False" ...
N=$1; shift
for o in "$@"; do
  DVS_DDP_OPTS="$o" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus $N --no-cpu --no-eager --no-train-big 2>/dev/null | tail -1 | O="$o" python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print('DDP_OPTS', repr(os.environ['O']), 'train ms/step', round(d['train_step']['ms_per_step'],3))"
done
