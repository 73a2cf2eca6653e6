for nq in 2 8 64 128; do
  python tools/batch_time.py 10000000 256 $nq 100 20 2>&1 | grep "ms/batch\|MISMATCH"
done
python tools/batch_time.py 12500000 64 8 100 20 2>&1 | grep "ms/batch\|MISMATCH"
python tools/batch_time.py 2500000 1024 8 100 20 2>&1 | grep "ms/batch\|MISMATCH"
python -m pytest tests/test_gpu_batched.py -x -q 2>&1 | tail -3
