for b in 13 14 15 16 17; do
python bench.py --batch $b --no-cpu --no-eager --no-train --no-train-big 2>/dev/null | tail -1 | B=$b python -c "import sys,json,os; d=json.loads(sys.stdin.read()); b=int(os.environ['B']); k=d['roofline']['kernel_ms']; print('B', b, 'kernel_ms', round(k,4), 'us per image', round(1000*k/b,2), 'waves', round(b*352/296,2))"
done
