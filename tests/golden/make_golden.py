#!/usr/bin/env python3
"""Regenerates tests/golden/ from the mounted reference (/root/reference).  Run in the build
container only; the GPU box has no reference and uses the committed files.

 - ref_host_1280x720.json : output of oracle/_ref/ref_dump, i.e. the REFERENCE'S OWN compiled
   loader / transform / camera / Halton / glm::inverse code and LuzCommon.h layouts
 - default.luz, default.luzbin.gz : the reference's default project (data asset, verbatim; the
   blob gzip'ed: it is two flat 1080x1080 textures + two cube meshes)
 - blue_noise_256.rgba : top-left 256x256 RGBA8 crop of assets/blue_noise.png (the shader
   addresses it with mod(fragCoord, size), light.frag:72-73, so any size is a valid input)
"""
import gzip, os, shutil, subprocess, sys, tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("LUZ_REFERENCE", "/root/reference")

def main():
    if not os.path.isdir(os.path.join(REF, "assets")):
        sys.exit("reference not mounted at %s" % REF)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
    with tempfile.TemporaryDirectory() as tmp:
        for (w, h) in ((1280, 720),):
            out = os.path.join(HERE, "ref_host_%dx%d.json" % (w, h))
            subprocess.check_call([exe, os.path.join(REF, "assets/default.luz"),
                                   os.path.join(REF, "assets/default.luzbin"), str(w), str(h), "40", out],
                                  cwd=tmp, stdout=subprocess.DEVNULL)
    shutil.copyfile(os.path.join(REF, "assets/default.luz"), os.path.join(HERE, "default.luz"))
    with open(os.path.join(REF, "assets/default.luzbin"), "rb") as f:
        blob = f.read()
    with open(os.path.join(HERE, "default.luzbin.gz"), "wb") as f:
        f.write(gzip.compress(blob, 9, mtime=0))
    from PIL import Image
    import numpy as np
    bn = np.array(Image.open(os.path.join(REF, "assets/blue_noise.png")).convert("RGBA"))
    bn[:256, :256].copy().tofile(os.path.join(HERE, "blue_noise_256.rgba"))
    print("golden fixtures written to", HERE)

if __name__ == "__main__":
    main()
