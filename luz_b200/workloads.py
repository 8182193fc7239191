"""The benchmark workloads of BASELINE.json / SURVEY.md section 8(d) as bench.py drives them: the synthetic project is
written to disk (scenes.write_project), loaded through the .luz loader of the host mirror, uploaded with
GPUScene::AddAssets, and stepped with RenderFrame (animated configs move their mesh nodes first).  bench.py times
these steps; tests/test_gpu_bench_parity.py checks the very same frames against the oracle.  Harness code."""
import json
import tempfile

import numpy as np

from . import host as H
from . import scenes


class Workload:
    def __init__(self, rt, config, variant=None, width=0, height=0, shadow_type=1, volumetric=0, tmp=None,
                 light_samples=None, ao_samples=None):
        self.rt, self.config, self.variant = rt, config, variant
        self.tmp = tmp or tempfile.mkdtemp(prefix="luzwork_")
        path, bin_path, cfg = scenes.write_project(config, self.tmp, variant)
        if shadow_type != 1 or volumetric:
            # SURVEY 8(f) rank 4 passes on the benchmark scene: shadow-map shadows instead of shadow rays and / or
            # volumetric lights (not the headline metric; reported in kernels_ms)
            with open(path) as f:
                doc = json.load(f)
            for sc in doc["scenes"].values():
                sc["shadowType"] = shadow_type

                def patch(nodes):
                    for n in nodes:
                        if n.get("type") == 7:
                            n["volumetricType"] = volumetric
                            n["shadowMapFar"] = 400.0
                        patch(n.get("children", []))
                patch(sc["nodes"])
            with open(path, "w") as f:
                json.dump(doc, f)
        if width:
            cfg["width"], cfg["height"] = width, height
        if light_samples is not None:  # experiments (profiles/): one kind of ray only
            cfg["light_samples"] = light_samples
        if ao_samples is not None:
            cfg["ao_samples"] = ao_samples
        self.cfg = cfg
        self.width, self.height = cfg["width"], cfg["height"]
        self.app = H.LuzHost(rt)
        self.app.load_project(path, bin_path)
        self.app.scene_settings(light_samples=cfg["light_samples"], ao_samples=cfg["ao_samples"])
        self.animate = cfg["animate"]
        self.frame = 0
        self._base = None

    def upload(self, blue_noise):
        """CreateImages + blue noise + AddAssets (BLAS builds): what Luz does once after loading a project."""
        self.app.set_extent(self.width, self.height, create_images=self.rt is not None)
        if self.rt is not None:
            self.rt.set_blue_noise(blue_noise)
        self.app.add_assets()
        if self.animate:
            n = self.app.mesh_node_count()
            base = [self.app.get_mesh_node_transform(i) for i in range(n)]
            self._base = (np.array([b[0] for b in base], np.float32), np.array([b[1] for b in base], np.float32))

    def move(self, frame):
        """Per-frame instance motion of the animated configs (C2: yaw, C5: translation)."""
        if not self.animate:
            return
        n = self.app.mesh_node_count()
        pos, rot = scenes.animate(self.config, frame, n, self._base[0], self._base[1])
        if self.config == "c2":
            self.app.set_mesh_node_transforms(1, rot=rot[1:])
        else:
            self.app.set_mesh_node_transforms(0, pos=pos)

    def step(self, first=False):
        """One frame.  The first frame builds the TLAS and produces the G-buffer on the device; animated configs
        re-run UpdateResources[GPU] (TLAS refit / rebuild) and the G-buffer producer every frame (moving instances
        change primary visibility; the producer is not part of the metric)."""
        if self.animate:
            self.move(self.frame)
            refit = H.FRAME_TLAS_REFIT if (self.animate == "refit" and not first) else 0
            self.app.render_frame(H.FRAME_OPAQUE | refit)
        else:
            self.app.render_frame(H.FRAME_OPAQUE if first else H.FRAME_NO_UPDATE)
        self.frame += 1
