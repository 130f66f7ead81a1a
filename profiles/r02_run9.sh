#!/bin/bash
DESC_BENCH_TRACE=1 python bench.py --workload cfg5 --steps 8 --warmup 3 --no-side --no-cpu 2>&1 | grep -E "^step|metric" | cut -c1-200
