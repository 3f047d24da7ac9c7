"""Development tool: host-side timing against the reference (profiles/r02_host_timings.txt)."""
import os, sys, time, tempfile, ctypes as C
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[os.path.join(ROOT,'yocto-gl_b200'),os.path.join(ROOT,'tests'),os.path.join(ROOT,'oracle')]
import refbind, scene_data
from ygl_b200 import lib
ref=refbind.Ref(); L=lib.load()
tmp=tempfile.mkdtemp()
def ours(p):
    h=C.c_void_p(); t=time.time(); rc=L.ygl_scene_load(str(p).encode(), C.byref(h)); dt=time.time()-t; L.ygl_loaded_scene_destroy(h); return dt
ref.lib.ref_scene_load.restype=C.c_void_p
def theirs(p):
    t=time.time(); h=ref.lib.ref_scene_load(str(p).encode()); dt=time.time()-t; ref.lib.ref_scene_destroy(C.c_void_p(h)); return dt
for name,data in [("features2",None),("shapes2",None),("shapes1",None),("shapes3",scene_data.DATA_V40),("features1",None),("materials4",None)]:
    p=scene_data.scene_file(name,tmp+("/v40" if data else ""),data)
    to=min(ours(p) for _ in range(5)); tr=min(theirs(p) for _ in range(5))
    print(f"{name:12s} ours {to*1000:7.1f} ms   reference {tr*1000:7.1f} ms")
p=scene_data.pool("shapes","bunny.ply")
print("bunny.ply as scene: ours %.1f ms reference %.1f ms"%(min(ours(p) for _ in range(3))*1000, min(theirs(p) for _ in range(3))*1000))
