#!/bin/bash
# DEVELOPER HARNESS ONLY (see mapcaller_b200/csrc/mc_device.h): compiles the stage bodies as plain C++ so
# their logic can be debugged against oracle/_ref on a box without a GPU.  The result is never loaded by
# the package, the tests or bench.py.
set -e
cd "$(dirname "$0")/../.."
g++ -O2 -g -std=c++17 -fPIC -shared -DMC_HOSTEMU -Wall -Wno-unused-function -Wno-unused-variable \
    -x c++ mapcaller_b200/csrc/mc_ctx.cu -x c++ mapcaller_b200/csrc/index.cpp mapcaller_b200/csrc/error.cpp \
    -o tools/hostemu/_build/libmc_hostemu.so -lpthread
