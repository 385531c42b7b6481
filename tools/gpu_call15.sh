#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/overlap_timeline.py $O/c15_timeline_graph_ov2.json wgrad_overlap=2 --graph > $O/c15_timeline_graph_ov2.txt 2>&1; head -30 $O/c15_timeline_graph_ov2.txt
timeout 300 python tools/overlap_timeline.py $O/c15_timeline_graph_ov0.json wgrad_overlap=0 --graph > $O/c15_timeline_graph_ov0.txt 2>&1; head -3 $O/c15_timeline_graph_ov0.txt
timeout 300 python tools/overlap_timeline.py $O/c15_timeline_eager_ov2.json wgrad_overlap=2 > $O/c15_timeline_eager_ov2.txt 2>&1; head -16 $O/c15_timeline_eager_ov2.txt
