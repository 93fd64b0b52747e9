# Build a compile-time variant of the product library into build/variants/NAME.so (for tools/ab_variants.sh).
# usage: bash tools/build_variant.sh NAME SOURCE.cu "-DFLAG=1 ..."   (SOURCE relative to rte_rrtmgp_b200/csrc)
set -e
NAME=$1; SRC=$2; FLAGS=$3
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CSRC=$ROOT/rte_rrtmgp_b200/csrc
mkdir -p $ROOT/build/variants $ROOT/build/vobj
OBJ=$ROOT/build/vobj/${NAME}_$(echo $SRC | tr / _).o
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --extended-lambda -Xcompiler -fPIC,-Wall,-Wno-unused-function \
  -I$ROOT/include -I$CSRC $FLAGS ${PTXAS_V:+-Xptxas -v} -c $CSRC/$SRC -o $OBJ
OTHERS=$(ls $ROOT/build/obj/*.o | grep -v "$(echo $SRC | tr / _).o")
nvcc -shared -arch=sm_100a -o $ROOT/build/variants/$NAME.so $OBJ $OTHERS -Xlinker -Bsymbolic -lcudart
echo built build/variants/$NAME.so
