# usage: build_variant.sh NAME "-DFLAG ..."  -> tnco_b200/libtnco_b200_NAME.so (only the 32_1 and 32_2 shapes rebuilt with the flags)
set -e
cd /root/repo/tnco_b200/csrc
NAME=$1; FLAGS=$2
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -ftz=true -Xcompiler -fPIC,-Wall,-Wno-unused-function -cudart static"
mkdir -p build_$NAME
for sh in 4_1 8_1 16_1 32_1 32_2 32_3 32_4; do
  T=${sh%_*}; W=${sh#*_}
  $NV $FLAGS -DTNB_INST_TILE=$T -DTNB_INST_WPL=$W -c -o build_$NAME/tnb_inst_$sh.o tnb_inst.cu &
done
$NV $FLAGS -c -o build_$NAME/tnb_engine.o tnb_engine.cu &
$NV $FLAGS -c -o build_$NAME/tnb_host.o tnb_host.cpp &
wait
$NV -shared -o ../libtnco_b200_$NAME.so build_$NAME/*.o
ls -la ../libtnco_b200_$NAME.so
