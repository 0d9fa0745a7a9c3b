"""GPU debug: where does GraphedDDPTrainStep (world 1) leave the eager train_step?  Runs both from identical state with
injected noise and prints, per sub-step, the loss and the largest parameter difference."""
import copy, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from util import build_model, load_arrays, rel_err
from mcgm_b200 import train_gan, ops

case = json.load(open(os.path.join(ROOT, "tests/golden/mcgan_golden.json")))["cases"]["tiny_cifar"]
arr = load_arrays("tiny_cifar")
m_e = build_model(case, device="cuda").to("cuda"); m_g = copy.deepcopy(m_e)
m_e.train(True); m_g.train(True)
img, label = arr["img"].cuda(), arr["label"].cuda()
zs = [z.cuda() for z in arr["zs"]]
opt_g = train_gan.make_optimizers(m_g, capturable=True)
red = {k: train_gan.GradAllReduce(getattr(m_g, k), overlap=False) for k in ("generator", "discriminator")}
S = train_gan.GraphedDDPTrainStep(m_g, opt_g, red, case["batch"], inject_noise=True)
opt_e = train_gan.make_optimizers(m_e, capturable=True)

def worst():
    a, b = m_e.state_dict(), m_g.state_dict()
    return max((rel_err(b[k].float().cpu(), a[k].float().cpu()), k) for k in a if a[k].is_floating_point())

def gworst(net):
    fa, fb = opt_e[net].flat.grad, opt_g[net].flat.grad
    return float((fa - fb).norm() / (fa.norm() + 1e-30))

print("start", worst())
S.img.copy_(img); S.label.copy_(label)
for it in range(2):
    for i in range(5):
        opt_e["discriminator"].zero_grad(); opt_e["generator"].zero_grad()
        le = train_gan.d_forward_backward(m_e, img, label, zs[i])
        S.z_d.copy_(zs[i]); S.graph_d.replay()
        print(f"it{it} D{i}: loss eager {float(le):.6f} graph {float(S.d_loss):.6f} grad diff {gworst('discriminator'):.2e} state {worst()}")
        opt_e["discriminator"].step(); opt_g["discriminator"].step()
        print(f"   after adam: {worst()}")
    opt_e["discriminator"].zero_grad(); opt_e["generator"].zero_grad()
    le = train_gan.g_forward_backward(m_e, label, zs[5])
    S.z_g.copy_(zs[5]); S.graph_g.replay()
    print(f"it{it} G: loss eager {float(le):.6f} graph {float(S.g_loss):.6f} grad diff {gworst('generator'):.2e} state {worst()}")
    opt_e["generator"].step(); opt_g["generator"].step()
    print(f"   after adam: {worst()}")
