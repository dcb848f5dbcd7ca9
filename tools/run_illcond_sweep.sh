for t in 1e7 1e9 1e11; do
  echo "== EKFVIO_ILLCOND=$t"; EKFVIO_ILLCOND=$t python tools/step_error_probe.py 0 100 2>&1 | tail -3
  EKFVIO_ILLCOND=$t python bench.py --skip-klt --skip-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['kernel_ms'], d['checks']['ldlt_filters'])"
done
