ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-micro --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
cat > /tmp/stem1.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from recnext_b200.model import RecNextStem, replace_batchnorm
m = replace_batchnorm(RecNextStem(3, 64).eval().cuda())
x = torch.randn(256, 3, 224, 224, device="cuda").bfloat16()
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    for _ in range(3): m(x)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:stem_kernel -c 1 -s 2 -o gpurun_out/r2_stem -f python /tmp/stem1.py > gpurun_out/ncu_stem.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ffn_tc_kernel -c 1 -s 2 -o gpurun_out/r2b_ffn_tc_c256 -f python tools/ffn_prof.py 256 256 14 512 3 > gpurun_out/ncu_ffn2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ffn_tc_kernel -c 1 -s 2 -o gpurun_out/r2b_ffn_tc_c64 -f python tools/ffn_prof.py 256 64 56 128 3 > gpurun_out/ncu_ffn3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ffn_tc_kernel -c 1 -s 2 -o gpurun_out/r2b_ffn_tc_c512 -f python tools/ffn_prof.py 256 512 7 1024 3 > gpurun_out/ncu_ffn4.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -6
