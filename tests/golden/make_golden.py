"""Regenerates the committed golden fixtures.  Needs /root/reference (run in the build container only):
imports the reference's pure-python ``flygym/anatomy.py`` (package __init__ bypassed because it needs dm_control)."""
import hashlib, importlib.util, json, sys, types
from pathlib import Path
import numpy as np
import yaml

REF = Path("/root/reference/src/flygym")
HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

pkg = types.ModuleType("flygym"); pkg.__path__ = [str(REF)]; sys.modules["flygym"] = pkg
utils = types.ModuleType("flygym.utils"); utils.__path__ = [str(REF / "utils")]; sys.modules["flygym.utils"] = utils
for name, path in (("flygym.utils.exceptions", REF / "utils/exceptions.py"), ("flygym.utils.math", REF / "utils/math.py"), ("flygym.anatomy", REF / "anatomy.py")):
    spec = importlib.util.spec_from_file_location(name, path); mod = importlib.util.module_from_spec(spec); sys.modules[name] = mod; spec.loader.exec_module(mod)
A = sys.modules["flygym.anatomy"]

full = A.Skeleton(joint_preset=A.JointPreset.ALL_POSSIBLE, axis_order=A.AxisOrder.DONTCARE)
segs = ["c_thorax"] + [d.child.name for d in full.iter_jointdofs("c_thorax") if d.axis == A.RotationAxis.PITCH]
sk = A.Skeleton(joint_preset=A.JointPreset.LEGS_ONLY, axis_order=A.AxisOrder.YAW_PITCH_ROLL)
dofs = list(sk.iter_jointdofs())
act = sk.get_actuated_dofs_from_preset(A.ActuatedDOFPreset.LEGS_ACTIVE_ONLY)
contact = [s.name for s in A.ContactBodiesPreset.LEGS_THORAX_ABDOMEN_HEAD.to_body_segments_list()]
pose = yaml.safe_load(open(REF / "assets/model/pose/neutral/yaw_pitch_roll.yaml"))["joint_angles"]
def neutral(name):
    if name in pose: return float(np.deg2rad(pose[name]))
    p, c, ax = name.split("-")
    mirror = f"{('l' + p[1:]) if p[0] == 'r' else p}-l{c[1:]}-{ax}"
    return float(np.deg2rad(pose.get(mirror, 0.0)))
out = dict(bodysegs_order=segs, jointdofs_legs_only_ypr=[d.name for d in dofs], actuated_legs_active_only=[d.name for d in act],
           contact_legs_thorax_abdomen_head=contact, neutral_angles_rad=[neutral(d.name) for d in dofs], legs=list(A.LEGS))
(HERE / "anatomy_orders.json").write_text(json.dumps(out, indent=1))

from flygym_b200 import retina as R
from oracle.retina_oracle import retina_oracle
idm = R.ommatidia_id_map(); pale = R.pale_mask(721)
img = np.random.default_rng(0).integers(0, 256, (1, 2, 512, 450, 3), dtype=np.uint8)
g = dict(id_map_sha256=hashlib.sha256(idm.tobytes()).hexdigest(), pale_mask_sha256=hashlib.sha256(np.packbits(pale).tobytes()).hexdigest(),
         oracle_seed0_sha256=hashlib.sha256(retina_oracle(img, idm, pale).tobytes()).hexdigest())
(HERE / "retina_golden.json").write_text(json.dumps(g, indent=1))

# ---- replay clip: run the reference's own MotionSnippet resampler (scipy only) and keep sample rows
demo = types.ModuleType("flygym_demo"); demo.__path__ = [str(REF.parent / "flygym_demo")]; sys.modules["flygym_demo"] = demo
sd = types.ModuleType("flygym_demo.spotlight_data"); sd.__path__ = [str(REF.parent / "flygym_demo/spotlight_data")]; sys.modules["flygym_demo.spotlight_data"] = sd
spec = importlib.util.spec_from_file_location("flygym_demo.spotlight_data.preprocessing", REF.parent / "flygym_demo/spotlight_data/preprocessing.py")
pre = importlib.util.module_from_spec(spec); spec.loader.exec_module(pre)
snip = pre.MotionSnippet(data_path=REF.parent / "flygym_demo/spotlight_data/assets/spotlight_behavior_clip.npz")
ref_angles = snip.get_joint_angles(1e-4, act)
rows = [0, 1, 777, 5000, 12345, 19999]
np.savez_compressed(HERE / "replay_golden.npz", rows=np.array(rows), angles=ref_angles[rows], shape=np.array(ref_angles.shape))
print("golden written", len(segs), len(dofs), len(act), len(contact), ref_angles.shape)
