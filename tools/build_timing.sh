#!/bin/sh
# developer build of the library with the clock64() timelines compiled in (never shipped: build/ is git-ignored)
set -e
cd "$(dirname "$0")/../se3-equi-graph-registration_b200/csrc"
mkdir -p ../../build/timing ../../build/timing_nowg ../../build/timing_solo
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DEGSPR_XB_TIMING -DEGSPR_TS_TIMING \
    -o ../../build/timing/libegspr_b200.so knn.cu csr.cu egnn_layer.cu egnn_edge_ts.cu egnn_node_ts.cu head.cu feature_match.cu egnn_backward.cu egnn_edge_bwd_tc.cu
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DEGSPR_XB_TIMING -DEGSPR_TS_TIMING -DEGSPR_XB_NO_WG \
    -o ../../build/timing_nowg/libegspr_b200.so knn.cu csr.cu egnn_layer.cu egnn_edge_ts.cu egnn_node_ts.cu head.cu feature_match.cu egnn_backward.cu egnn_edge_bwd_tc.cu
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DEGSPR_XB_TIMING -DEGSPR_TS_TIMING -DEGSPR_XB_SOLO \
    -o ../../build/timing_solo/libegspr_b200.so knn.cu csr.cu egnn_layer.cu egnn_edge_ts.cu egnn_node_ts.cu head.cu feature_match.cu egnn_backward.cu egnn_edge_bwd_tc.cu
