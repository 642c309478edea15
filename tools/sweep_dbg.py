"""Per-call timing of 1024-utterance generate_batch calls on each rank (run under torchrun or alone).

Used to separate first-call costs (capacity growth, pinned staging) from the steady state of the sweep:
PTTS_B200_E2E_TRACE=1 python tools/sweep_dbg.py
"""
import os, sys, time, numpy as np, torch
sys.path.insert(0,'.')
import bench
from __graft_entry__ import load_package
rank=int(os.environ.get("LOCAL_RANK","0"))
torch.cuda.set_device(rank)
pkg=load_package()
md=pkg.synth.model_dir(0,125)
W=pkg.synth.read_safetensors(md/"tts_b6369a24.safetensors")
voice=pkg.synth.read_safetensors(md/"embeddings"/"alba.safetensors")["audio_prompt"].reshape(-1,1024)
eng=pkg.Engine(W,rank)
toks=bench.tokenize_prompts(pkg, md, pkg.synth.make_prompts(1024, seed=777+rank))
seeds=np.arange(1024,dtype=np.int64)+9000
params=pkg.default_params(lsd_steps=1,temp=0.7,eos_enabled=0)
audio=torch.empty((1024,100*1920),dtype=torch.float32).pin_memory().numpy()
os.environ["PTTS_B200_E2E_TRACE"]="1"
for i in range(4):
    t=time.perf_counter(); eng.generate_batch(toks,voice,100,seeds,None,params,True,audio); print(f"rank {rank} call {i}: {1e3*(time.perf_counter()-t):.1f} ms", file=sys.stderr)
