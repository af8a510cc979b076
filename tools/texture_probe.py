import sys, numpy as np, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rustracer_b200 import Scene, scenes
from rustracer_b200.device import Device
from oracle import binding as ob
d='/tmp/texscene_gpu'
dev=Device(0)
for name,kw in [("path",{}),("lens",dict(lens=True)),("whitted",dict(integrator='Integrator "whitted" "integer maxdepth" [4]')),("direct",dict(integrator='Integrator "directlighting" "integer maxdepth" [3] "string strategy" "one"'))]:
    sc=Scene.from_string(scenes.balls_textured(d, xres=96, yres=72, spp=8, **kw), search_dir=d)
    dev.upload(sc)
    o=ob.OracleScene(sc.ir_ptr)
    rd=sc.render_desc(); rd.seed=7
    rng=np.random.default_rng(1); sb=list(rd.sample_bounds); n=4000
    pix=np.stack([rng.integers(sb[0],sb[2],n),rng.integers(sb[1],sb[3],n),rng.integers(0,rd.spp,n)],1).astype(np.int32)
    ref,_=o.li_samples(pix,seed=7); got=dev.li_samples(rd,pix)
    tol=1e-4*np.maximum(np.abs(ref),1e-3)+1e-6
    bad=(np.abs(got-ref)>tol).any(1)
    rel=np.abs(got-ref)/np.maximum(np.abs(ref),1e-3)
    print(name,"bad frac",bad.mean(),"max rel",rel.max(), "nan", np.isnan(got).sum())
    for k in np.where(bad)[0][:5]: print("   ",pix[k],ref[k],got[k])
    st=dev.render(rd); rgb=dev.resolve_film()
    film_ref,rgb_ref,ost=o.render(sampler_kind=1,seed=7)
    print("   image rel", np.abs(rgb-rgb_ref).sum()/np.abs(rgb_ref).sum(), st.camera_rays==ost.camera_rays, st.regular_rays, ost.regular_rays, st.shadow_rays, ost.shadow_rays)
