import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d["roofline"]; k=d["kernels_ms"]
        print("%-28s ms/step %.3f light %.3f taa %.3f tlas %.3f  nodes/ray %.3f tris/ray %.3f inst/ray %.3f  Grays/s %.2f occl %.3f" % (f.split('/')[-1], d["ms_per_step"], k["light"], k["taa"], k["tlas"], r["nodes_per_ray"], r["tris_per_ray"], r["instances_per_ray"], d["value"]/1e3, r.get("occluded_fraction",-1)))
    except Exception as e:
        print(f, "ERR", e)
