#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python tools/bench_graph_error.py 0 10 11 12 13 5 > gpurun_out/r2r_bench_graph_error.log 2>&1; cat gpurun_out/r2r_bench_graph_error.log
