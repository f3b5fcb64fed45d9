mkdir -p gpurun_out; rm -f gpurun_out/multirank_parity.jsonl
timeout 300 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu 2>&1 | tail -5
cp gpurun_out/multirank_parity.jsonl gpurun_out/r02_multirank_parity_2gpu.jsonl; cat gpurun_out/r02_multirank_parity_2gpu.jsonl | cut -c 1-300
bash tools/jobs/job_scale.sh 2
