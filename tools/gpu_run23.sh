mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv gpurun_out/*.gz
for cfg in "0 0" "1 0" "1 1" "0 1"; do set -- $cfg; echo "channels_last=$1 native_bn=$2"; WS3D_TRAIN_CHANNELS_LAST=$1 WS3D_NATIVE_BN=$2 timeout 300 python tools/train_rpn_bench.py --graph 0 2>&1 | tail -1 | cut -c1-200; done
WS3D_NATIVE_BN=0 timeout 300 python tools/train_profile.py 2>&1 | tail -34 | cut -c1-130 | head -16
