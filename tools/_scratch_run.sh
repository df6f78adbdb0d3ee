# scratch command file for `gpurun -- bash tools/_scratch_run.sh` (overwritten per experiment)
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 2>&1 | tail -3
timeout 900 python bench.py
