#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/bin/probe_graph_chain > gpurun_out/graph_chain.txt 2>&1; cat gpurun_out/graph_chain.txt
B200Q_LIB=$PWD/convnet_quantization_b200/libb200q_dev.so timeout 600 python scripts/graph_breakdown.py > gpurun_out/graph_breakdown.json 2> gpurun_out/graph_breakdown.log; echo "breakdown exit=$?"; cat gpurun_out/graph_breakdown.log
