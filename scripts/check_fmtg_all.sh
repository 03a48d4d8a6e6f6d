#!/bin/bash
# csrc/xm_fmtg.h against the C library's "%g" for EVERY 32-bit float (about five minutes on eight cores)
set -e
cd "$(dirname "$0")/.."
g++ -O2 -std=c++17 -pthread -o /tmp/fmtg_check tests/emu/fmtg_check.cpp
/tmp/fmtg_check 1 "${1:-$(nproc)}"
