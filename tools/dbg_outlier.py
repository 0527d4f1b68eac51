import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from graspit_b200.engine import B2C_INPUT_AA_EIGEN, Engine, SceneBinding
from graspit_b200.formats.scenepack import load_pack
from oracle import ref
from oracle.port import fk as ofk, transf_np as T
from tests.common import OracleScene, object_body, sample_aa_states
npre, seed = 200000, 777
sc = load_pack("barrett_mug"); obj = object_body(sc); r = sc.robots[0]
eng = Engine(0); sb = SceneBinding(eng, sc)
osc = OracleScene(sc, nthreads=ref.hardware_threads())
rng = np.random.default_rng(seed)
pos, _ = sample_aa_states(sc, npre, seed=seed, strata=("far", "near", "close"))
states = np.concatenate([pos, rng.uniform(-4, 4, size=(npre, r.eg.shape[0]))], 1).astype(np.float32)
s64 = states.astype(np.float64)
base = ofk.aa_state_to_base(r, sc.body_tran[obj], s64[:, :6])
raw64 = np.concatenate([base[0], base[1], ofk.eigen_to_dof(r, s64[:, 6:])], 1)
lo, eo, _, order, tf7 = osc.eval(0, obj, raw64, mode=0)
res = eng.eval_batch(sb.robot_ids[0], sb.body_ids[obj], states, input_mode=B2C_INPUT_AA_EIGEN, ref_tran=sc.body_tran[obj])
m = (lo == 1)
rel = np.where(m, np.abs(res["energy"] - eo) / np.maximum(np.abs(eo), 1e-9), 0)
p = int(np.argmax(rel)); print("pose", p, "rel", rel[p], "engine", res["energy"][p], "oracle", eo[p])
w = osc.w
for j, b in enumerate(order):
    w.set_transform(osc.ids[b], tf7[p, j, :4], tf7[p, j, 4:])
oq, ot = sc.body_tran[obj]
eng.set_body_transform(sb.body_ids[obj], oq, ot)
for v in r.vcontacts:
    j = order.index(v.body)
    tl = T.make(tf7[p, j, :4][None], tf7[p, j, 4:][None])
    pw = T.apply(tl, np.array(v.loc)[None])[0]
    cnw = T.rotate(tl, np.array(v.normal)[None])[0]
    do, cpo, cno = w.point_distance(osc.ids[obj], pw)
    de, cpe, cne = eng.point_to_body_distance(sb.body_ids[obj], pw)
    to = do + 50 * (1 - cnw @ ((cpo - pw) / do)); te = de + 50 * (1 - cnw @ ((cpe - pw) / de))
    flag = "  <<<" if abs(to - te) > 1e-3 else ""
    print("contact d_oracle %.7f d_engine(point_kernel) %.7f  term_o %.5f term_e %.5f |cp diff| %.4g%s" % (do, de, to, te, np.linalg.norm(cpo - cpe), flag))
