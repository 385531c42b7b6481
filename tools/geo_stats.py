import sys; sys.path.insert(0,'.'); sys.path.insert(0,'complex-yolov4-pytorch_b200')
import numpy as np, torch
from cy4 import geometry as cg, synth
from oracle import geometry as og
g=np.load('tests/golden/rgiou_pairs.npz')
def run(pred,tgt,giou=True):
    p=torch.tensor(pred,device='cuda',requires_grad=True); t=torch.tensor(tgt,device='cuda')
    i,tm=cg.rgiou_pairs(p,t,giou); tm.sum().backward()
    return i.cpu().numpy(), tm.detach().cpu().numpy(), p.grad.cpu().numpy()
i,t,gr=run(g['pred'],g['tgt'])
d=np.abs(i-g['iou']); print("golden: iou maxdiff",d.max(),"n>1e-4",(d>1e-4).sum(),"n>1e-5",(d>1e-5).sum(),"biteq",(i==g['iou']).mean())
d=np.abs(t-g['term']); print("golden: term maxdiff",d.max(),"n>1e-4",(d>1e-4).sum())
gd=np.abs(gr-g['grad'])/(np.abs(g['grad'])+1e-2); print("golden grad relmax",gd.max())
oi,ot,ogr=og.rgiou_pairs(g['pred'],g['tgt'],True,True)
d=np.abs(oi-g['iou']); print("oracle vs golden: iou maxdiff",d.max(),"biteq",(oi==g['iou']).mean())
n=len(g['shapely_iou']); i2,t2,g2=run(g['pred'][:n],g['tgt'][:n],False)
print("shapely iou maxdiff",np.abs(i2-g['shapely_iou']).max(),"grad",np.abs(g2-g['shapely_grad']).max())
for n,seed,dj in [(127,1,0.1),(128,2,0.0),(100000,7,0.01)]:
    pred,tgt=synth.make_pairs(n,seed=seed,disjoint_frac=dj)
    i,t,gr=run(pred,tgt); oi,ot,ogr=og.rgiou_pairs(pred,tgt,True,True)
    ei,et=og.rgiou_pairs_exact64(pred,tgt)
    d=np.abs(i-oi); bad=d>1e-4
    print(n,"vs oracle iou max",d.max(),"n>1e-4",bad.sum(),"biteq",(i==oi).mean(),"term max",np.abs(t-ot).max(),"nan",np.isnan(i).sum(),np.isnan(oi).sum())
    if bad.any():
        k=np.argmax(d); print("  worst",k,pred[k],tgt[k],i[k],oi[k],ei[k])
        print("  cuda-vs-exact on bad",np.abs(i[bad]-ei[bad]).max(),"oracle-vs-exact",np.abs(oi[bad]-ei[bad]).max())
    print("  grad rel max",(np.abs(gr-ogr)/(np.abs(ogr)+1e-2)).max())
# identical boxes
pred,tgt=synth.make_pairs(100000,seed=7)
i,t,_=run(tgt,tgt); oi,ot=og.rgiou_pairs(tgt,tgt,True)
m=~(np.isnan(i)|np.isnan(oi))
print("identical: cuda nan",np.isnan(i).sum(),"oracle nan",np.isnan(oi).sum(),"maxdiff",np.abs(i[m]-oi[m]).max(),"iou range",np.nanmin(i),np.nanmax(i),np.nanmin(oi),np.nanmax(oi), "nan same",(np.isnan(i)==np.isnan(oi)).mean())
