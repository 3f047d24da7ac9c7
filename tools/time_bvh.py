"""Development tool: host-side timing against the reference (profiles/r02_host_timings.txt)."""
import os, sys, time, tempfile, os, ctypes as C
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[os.path.join(ROOT,'yocto-gl_b200'),os.path.join(ROOT,'tests'),os.path.join(ROOT,'oracle')]
import refbind, scene_data
from ygl_b200 import lib, scenes
ref=refbind.Ref(); L=lib.load()
tmp=tempfile.mkdtemp()
pool=os.path.join(scene_data.DATA,"pool")
cases={"c2 bunny":scenes.bunny_file_scene(tmp,pool),"c5 hairball":scenes.hairball_file_scene(tmp,pool),"c3":scenes.instanced_spheres(10)}
for name,sc in cases.items():
    d=sc.desc()
    for hq in (False,True):
        best=1e9
        for _ in range(3):
            h=C.c_void_p(); t=time.time(); L.ygl_bvh_build(C.byref(d),int(hq),C.byref(h)); best=min(best,time.time()-t); L.ygl_bvh_destroy(h)
        rs=ref.scene(sc); bt=1e9
        for _ in range(2):
            t=time.time(); b=ref.lib.ref_bvh_build(rs.h,int(hq)); bt=min(bt,time.time()-t)
        mine=lib.Bvh(sc,hq); ok=True
        for shape in [-1]+list(range(len(sc.shapes))):
            n_ref,p_ref=rs.bvh_tree(shape,hq); n_my,p_my=mine.tree(shape)
            ok = ok and n_ref.tobytes()==n_my.tobytes() and p_ref.tobytes()==p_my.tobytes()
        print(f"{name:12s} highquality={hq}: ours {best*1000:7.2f} ms  reference {bt*1000:7.2f} ms  identical={ok}")
