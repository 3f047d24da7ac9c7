"""Development tool: byte-level mutation fuzzing of every file reader of the library (scene JSON, glTF, OBJ scenes, PLY / OBJ /
STL shapes, PNG / JPEG / HDR / EXR textures) under AddressSanitizer + UBSan. Seeds come from the generators of the differential
tests; each mutant must be loaded or refused without a sanitizer report. Build the harness first:
  g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -ffp-contract=off tools/loader_fuzz/harness.cpp \
      yocto-gl_b200/csrc/ygl_sceneio.cpp yocto-gl_b200/csrc/ygl_imageio.cpp -o $FUZZ_WORK/harness -lz -lpthread
usage: ASAN_OPTIONS=detect_leaks=0 python tools/loader_fuzz/mutate.py <seed> <files>"""
import os, sys, random, subprocess, pathlib, io, json, struct, shutil
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0]=[os.path.join(ROOT,'tests'),os.path.join(ROOT,'yocto-gl_b200'),os.path.join(ROOT,'oracle')]
os.environ["OPENCV_IO_ENABLE_OPENEXR"]="1"
import numpy as np
import test_texture_fuzz as TT, test_shape_fuzz as TS, test_gltf_fuzz as TG, test_sceneio_fuzz as TJ
import cv2
from PIL import Image
seed=int(sys.argv[1]); rng=random.Random(seed)
WORK=os.environ.get("FUZZ_WORK","/tmp/ygl_loader_fuzz")
root=pathlib.Path(WORK)/("work%d"%seed); shutil.rmtree(root,ignore_errors=True); root.mkdir(parents=True)
# seeds
seeds=[]
for k in range(6):
    ct,dp=TT.VALID[rng.randrange(len(TT.VALID))]
    seeds.append(("png",TT.make_png(rng,rng.randint(1,20),rng.randint(1,20),ct,dp,rng.randrange(2))))
for k,(d,w) in enumerate(TT._jpeg_cases(rng,8)): seeds.append(("jpg",d))
for k in range(4): seeds.append(("hdr",TT.make_hdr(rng,rng.randint(1,40),rng.randint(1,5),rle=rng.random()<0.5)))
for comp in (0,1,2,3,4,4):
    px=np.random.default_rng(seed+comp).normal(0,1,(rng.randint(1,40),rng.randint(1,40),3)).astype(np.float32)
    ok,d=cv2.imencode(".exr",px,[cv2.IMWRITE_EXR_COMPRESSION,comp,cv2.IMWRITE_EXR_TYPE,rng.choice([1,2])]); seeds.append(("exr",d.tobytes()))
for k in range(5): seeds.append(("ply",TS.make_ply(rng)))
for k in range(3): seeds.append(("objshape",TS.make_obj(rng)))
gl=root/"gl"; gl.mkdir()
for k in range(5):
    p=TG.make_gltf(rng,gl,f"g{k}"); seeds.append(("gltf:"+str(p),p.read_bytes()))
ob=root/"ob"; ob.mkdir()
for k in range(4):
    p=TS.make_obj_scene(rng,ob,f"o{k}"); seeds.append(("objscene:"+str(p),p.read_bytes()))
(root/"shapes").mkdir()
from test_sceneio import _write_tri_ply
_write_tri_ply(root/"shapes"/"tri.ply")
for k in range(5): seeds.append(("json",TJ.make_document(rng)))
stl=b"x".ljust(80,b"\0")+struct.pack("<I",3)+b"".join(struct.pack("<12fH",*[rng.random() for _ in range(12)],0) for _ in range(3)); seeds.append(("stl",stl))
def mutate(data):
    b=bytearray(data)
    r=rng.random()
    n=max(1,int(len(b)*rng.choice([0.001,0.01,0.05])))
    if r<0.5:
        for _ in range(n): b[rng.randrange(len(b))]=rng.randrange(256)
    elif r<0.7: b=b[:rng.randrange(len(b)+1)]
    elif r<0.85:
        i=rng.randrange(len(b)); b[i:i+rng.randint(1,8)]=bytes(rng.choice([0,255,127,128]) for _ in range(rng.randint(1,8)))
    else:
        i=rng.randrange(len(b)); b[i:i]=bytes(rng.randrange(256) for _ in range(rng.randint(1,16)))
    return bytes(b)
files=[]
N=int(sys.argv[2])
for it in range(N):
    kind,data=seeds[rng.randrange(len(seeds))]
    m=mutate(data) if rng.random()<0.95 else data
    if kind.startswith("gltf:") or kind.startswith("objscene:"):
        src=pathlib.Path(kind.split(":",1)[1]); dst=src.parent/(f"m{it}"+src.suffix); dst.write_bytes(m); files.append(str(dst))
    elif kind=="json": p=root/f"m{it}.json"; p.write_bytes(m); files.append(str(p))
    elif kind=="objshape": p=root/f"m{it}.obj"; p.write_bytes(m); files.append(str(p))
    elif kind=="ply": p=root/f"m{it}.ply"; p.write_bytes(m); files.append(str(p))
    elif kind=="stl":
        (root/"shapes"/f"m{it}.stl").write_bytes(m); p=root/f"s{it}.json"; p.write_text(json.dumps({"asset":{"version":"4.2"},"shapes":[{"uri":f"shapes/m{it}.stl"}]})); files.append(str(p))
    else: p=root/f"m{it}.{kind}"; p.write_bytes(m); files.append(str(p))
bad=0
for i in range(0,len(files),40):
    r=subprocess.run([os.path.join(WORK,"harness")]+files[i:i+40],stdout=subprocess.PIPE,stderr=subprocess.PIPE,text=True,timeout=600)
    if r.returncode!=0:
        # find the culprit
        for f in files[i:i+40]:
            r1=subprocess.run([os.path.join(WORK,"harness"),f],stdout=subprocess.PIPE,stderr=subprocess.PIPE,text=True,timeout=120)
            if r1.returncode!=0:
                bad+=1; print("CRASH",f); print("\n".join(l for l in r1.stderr.splitlines() if "ERROR" in l or "#0" in l or "#1 " in l or "#2 " in l or "runtime error" in l)[:900])
print("files",len(files),"crashes",bad)
