import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mobile_env_gan_b200 as mbe
E=65536; R=4
envs=[mbe.make("mobile-medium-central-v0", num_envs=E, autoreset=True, env_offset=r*E) for r in range(R)]
g=torch.Generator(device="cuda").manual_seed(0)
for e in envs:
    e.reset(); e.actions.copy_(torch.randint(0,5,(E,15),generator=g,device="cuda",dtype=torch.int32))
torch.cuda.synchronize()
def bench(nstreams, steps=4096):
    streams=[torch.cuda.Stream() for _ in range(nstreams)]
    graphs=[]
    per=steps//nstreams
    chunk=128
    for si,s in enumerate(streams):
        mine=[envs[i] for i in range(R) if i % nstreams == si]
        with torch.cuda.stream(s):
            for i in range(8): mine[i%len(mine)].step(mine[i%len(mine)].actions)
            torch.cuda.synchronize()
            gr=torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                for i in range(chunk): mine[i%len(mine)].step(mine[i%len(mine)].actions)
        graphs.append(gr)
    torch.cuda.synchronize()
    for rep in range(3):
        for gr,s in zip(graphs,streams):
            with torch.cuda.stream(s): gr.replay()
    torch.cuda.synchronize()
    ev0=torch.cuda.Event(enable_timing=True); ev1=torch.cuda.Event(enable_timing=True)
    ev0.record()
    for s in streams: s.wait_event(ev0)
    for rep in range(per//chunk):
        for gr,s in zip(graphs,streams):
            with torch.cuda.stream(s): gr.replay()
    cur=torch.cuda.current_stream()
    for s in streams: cur.wait_stream(s)
    ev1.record()
    torch.cuda.synchronize()
    ms=ev0.elapsed_time(ev1); n=(per//chunk)*chunk*nstreams
    print(nstreams, "streams:", round(ms*1e3/n,2), "us/step", f"{E*n/(ms*1e-3):.3e} env-steps/s")
bench(1); bench(2); bench(4); bench(1)
