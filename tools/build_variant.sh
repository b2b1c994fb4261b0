#!/bin/bash
# Builds a variant of liblrag.so with extra -D flags into legal_rag_b200/variants/<name>.so (A/B runs: LRAG_LIB_PATH=...).
# usage: tools/build_variant.sh <name> <nvcc flags...>
NAME=$1; shift
mkdir -p legal_rag_b200/variants
cp legal_rag_b200/liblrag.so /tmp/liblrag_keep.so 2>/dev/null
LRAG_NVCC_EXTRA="$*" python -m legal_rag_b200.build --force > /dev/null && cp legal_rag_b200/liblrag.so legal_rag_b200/variants/$NAME.so
echo "built variants/$NAME.so ($*)"
