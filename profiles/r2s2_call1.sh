set -x
nvidia-smi -L
python profiles/far_mesh_parity.py 8 > gpurun_out/r2s2_far_mesh.jsonl 2> gpurun_out/r2s2_far_mesh.err; echo rc=$?
tail -3 gpurun_out/r2s2_far_mesh.err; cat gpurun_out/r2s2_far_mesh.jsonl
python profiles/run_configs.py hexq1 quadnitsche > gpurun_out/r2s2_cfg_staged.jsonl 2>gpurun_out/r2s2_cfg_staged.err; echo rc=$?
CQ_VARIANT=1 python profiles/run_configs.py hexq1 quadnitsche > gpurun_out/r2s2_cfg_direct.jsonl 2>gpurun_out/r2s2_cfg_direct.err; echo rc=$?
cat gpurun_out/r2s2_cfg_staged.jsonl gpurun_out/r2s2_cfg_direct.jsonl | cut -c1-200
