import sys, json, numpy as np, torch, torch.nn.functional as F
sys.path[:0]=['/root/repo','/root/repo/tests','/root/repo/efficient-segmentation-networks_b200']
from conftest import spec_state_dict
from oracle import fixture, nets
spec=json.load(open('/root/repo/tests/golden/state_dict_spec.json'))
g=np.load('/root/repo/tests/golden/ERFNet.npz')
def rel(a,b): return ((a.double()-b.double()).norm()/b.double().norm()).item()
for dev in ('cpu','cuda'):
    sd={k:(v.to(dev).requires_grad_(True) if v.is_floating_point() else v.to(dev)) for k,v in spec_state_dict(spec,'ERFNet').items()}
    x=fixture.make_input(2,64,128).to(dev); lab=fixture.make_labels(2,64,128,19).to(dev)
    y=nets.forward('ERFNet', sd, x, train=True)
    l=F.cross_entropy(y, lab, torch.tensor(fixture.CLASS_WEIGHTS, device=dev), ignore_index=255)
    l.backward()
    for key in g.files:
        if key.startswith('train_2x64x128_grad::'):
            k=key.split('::')[1]
            print(dev, 'oracle fp32 vs fp64 golden', k, rel(sd[k].grad.cpu(), torch.from_numpy(g[key])))
