mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused20.py -x -q -m gpu 2>&1 | tail -3
{
for sh in 8,8,2,1 8,8,2,0 4,4,4,1 16,8,1,1; do
P4B_AA3_SHAPE=$sh python tools/sweep_aa.py --want -32627327.455509827
done
P4B_AA3_SHAPE=8,8,2,1 python tools/sweep_aa.py --cfg 4
python tools/sweep_dna.py --variants 16,13,10 --lean
} 2>&1 | grep -v "^$\|^JSON" | tee gpurun_out/sweep_r2n.txt
