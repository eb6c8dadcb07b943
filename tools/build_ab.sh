#!/bin/sh
# developer A/B build: tools/build_ab.sh <name> [-DFLAG ...] -> build/<name>/libegspr_b200.so (git-ignored; tools/time_layer.py <name>)
set -e
name=$1; shift
cd "$(dirname "$0")/../se3-equi-graph-registration_b200/csrc"
mkdir -p ../../build/$name
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared "$@" \
    -o ../../build/$name/libegspr_b200.so knn.cu csr.cu egnn_layer.cu egnn_edge_ts.cu egnn_node_ts.cu head.cu feature_match.cu egnn_backward.cu egnn_edge_bwd_tc.cu
