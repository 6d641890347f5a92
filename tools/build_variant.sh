#!/bin/bash
# builds a variant of the library with extra nvcc flags into cooperative-search_b200/csrc/variants/<name>.so
# usage: tools/build_variant.sh name "-DCS_MAP_MIN_CTAS=14 -DCS_MAP_U=2"
set -e
name=$1; flags=$2
mkdir -p cooperative-search_b200/csrc/variants
CS_NVCC_EXTRA="$flags" python - <<PY
import importlib
b = importlib.import_module('cooperative-search_b200.build')
print(b.build_library(force=True))
PY
cp cooperative-search_b200/csrc/libcoopsearch.so cooperative-search_b200/csrc/variants/$name.so
