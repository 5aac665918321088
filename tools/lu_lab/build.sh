#!/bin/bash
# build lu_lab variants: build.sh <name> <header> [extra nvcc flags]
set -e
cd "$(dirname "$0")"
name=$1; hdr=$2; shift 2
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -I../../include -I../../jitr_b200/csrc \
  -DLU_HEADER="\"$hdr\"" "$@" lu_lab.cu -o bin/lab_$name
