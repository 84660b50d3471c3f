# tile-width experiments on the tuning library (shallow-K layers are epilogue-bound: are two narrow CTAs per SM faster?)
mkdir -p gpurun_out
export TORTTO_B200_LIB=tuning
for bn in 0 64 128; do
  echo "== TTB_FORCE_BN=$bn preact_resnet18 tf32"; TTB_FORCE_BN=$bn timeout 200 python scripts/per_layer.py preact_resnet18 | grep -v "^ttb_"
done > gpurun_out/force_bn_r18.txt 2>&1
for bn in 0 64 128; do
  echo "== TTB_FORCE_BN=$bn standard_resnet50 bf16"; TTB_FORCE_BN=$bn timeout 300 python scripts/per_layer.py standard_resnet50 | grep -v "^ttb_"
done > gpurun_out/force_bn_r50.txt 2>&1
tail -5 gpurun_out/force_bn_r18.txt gpurun_out/force_bn_r50.txt
