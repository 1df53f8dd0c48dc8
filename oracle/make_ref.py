"""Recipe: stage the UNMODIFIED reference files of the hot path under oracle/_ref/ so they travel to the GPU box.

    python oracle/make_ref.py            (authoring container only: needs /root/reference)

oracle/_ref/ is git-ignored (never committed -- the repository holds no reference source) but NOT gpurun-ignored, so
`gpurun` ships it with the snapshot, exactly like the in-tree .so.  On the box it is what
  * tests/test_patch_gpu.py builds GwcNet / PSMNet / StereoBase sub-graphs from (the reference's own classes + unchanged YAML),
  * bench.py --impl reference and the cpu_baseline leg time (kind "reference"),
  * bench.py's GPU comparators run (the reference on cuDNN fp32; its Triton gwc kernel).
Files are byte copies; a MANIFEST with sha256 sums is written next to them so a stale copy is detectable.
TEST / MEASUREMENT INFRASTRUCTURE ONLY: nothing under openstereo_b200/ reads oracle/_ref.
"""
import glob
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("OPENSTEREO_REFERENCE_SRC", "/root/reference")

# directories copied whole (python files only) and single files; relative to the reference root
DIRS = [
    "stereo/modeling/cost_volume", "stereo/modeling/disp_pred", "stereo/modeling/common", "stereo/modeling/disp_refinement",
    "stereo/modeling/backbones", "stereo/modeling/models/gwcnet", "stereo/modeling/models/psmnet",
    "stereo/modeling/models/stereobase", "stereo/modeling/models/igev", "stereo/modeling/models/lightstereo",
]
FILES = [
    "stereo/modeling/models/fast_foundationstereo/core/submodule.py",      # Triton gwc kernel: the existing-GPU-kernel comparator
    "stereo/modeling/models/igevpp/submodule.py", "stereo/modeling/models/casnet/submodule.py",
    "stereo/evaluation/metric_per_image.py", "tools/measure.py",
    "cfgs/gwcnet/gwcnet_sceneflow.yaml", "cfgs/psmnet/psmnet_sceneflow.yaml", "cfgs/stereobase/stereobase_sceneflow.yaml",
    "cfgs/lightstereo/lightstereo_s_sceneflow.yaml", "cfgs/igev/igev_sceneflow_amp.yaml",
]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def make(verbose=False):
    if not os.path.isdir(os.path.join(SRC, "stereo", "modeling")):
        raise RuntimeError("reference tree not found at %s" % SRC)
    wanted = list(FILES)
    for d in DIRS:
        wanted += sorted(os.path.relpath(p, SRC) for p in glob.glob(os.path.join(SRC, d, "*.py")))
    manifest = {}
    for rel in wanted:
        src = os.path.join(SRC, rel)
        if not os.path.exists(src):
            raise RuntimeError("missing reference file %s" % rel)
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and _sha(dst) == _sha(src)):
            shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
    stale = [os.path.relpath(os.path.join(r, f), DEST) for r, _, fs in os.walk(DEST) for f in fs
             if os.path.relpath(os.path.join(r, f), DEST) not in manifest and f != "MANIFEST.json"]
    for rel in stale:
        os.remove(os.path.join(DEST, rel))
    head = ""
    try:
        with open(os.path.join(SRC, ".git", "HEAD")) as f:
            head = f.read().strip()
    except OSError:
        pass
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "git_head": head, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print("oracle/_ref: %d files (%d bytes)" % (len(manifest), sum(os.path.getsize(os.path.join(DEST, r)) for r in manifest)))
    return DEST


if __name__ == "__main__":
    make(verbose=True)
    sys.exit(0)
