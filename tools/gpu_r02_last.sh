#!/bin/bash
# last GPU check of the round: entry points behind the exception guard (batched host call, several contexts, one-call
# combiner, C host program)
timeout 150 python -m pytest tests/test_gpu_parity.py tests/test_c_host.py -x -q -m gpu -k "concurrent or batch_matches or several_contexts or many_units or c_host" 2>&1 | tail -2
