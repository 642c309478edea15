import sys, time, numpy as np, torch
sys.path.insert(0,'.')
import bench
from __graft_entry__ import load_package
pkg=load_package()
md=pkg.synth.model_dir(0,125)
W=pkg.synth.read_safetensors(md/"tts_b6369a24.safetensors")
voice=pkg.synth.read_safetensors(md/"embeddings"/"alba.safetensors")["audio_prompt"].reshape(-1,1024)
eng=pkg.Engine(W,0)
toks=bench.tokenize_prompts(pkg, md, pkg.synth.make_prompts(256, seed=123))
seeds=np.arange(256,dtype=np.int64)+123
params=pkg.default_params(lsd_steps=1,temp=0.7,eos_enabled=0)
audio=torch.empty((256,100*1920),dtype=torch.float32).pin_memory().numpy()
for i in range(4):
    t=time.perf_counter(); eng.generate_batch(toks,voice,100,seeds,None,params,True,audio); print("python wall ms", (time.perf_counter()-t)*1e3, file=sys.stderr)
