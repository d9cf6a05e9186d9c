import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from model_based_pde_control_b200 import KSVecEnv
for N, L, J, B in ((64, 22.0, 4, 9), (128, 44.0, 4, 9), (128, 44.0, 4, 8), (256, 88.0, 8, 4)):
    rng = np.random.default_rng(1)
    env = KSVecEnv(B, dict(N=N, L=L, dt=0.025, cfg_steps=10), Xi=[k / J for k in range(J)], solver="etdrk4")
    u0 = rng.uniform(-1, 1, (B, N))
    acts = torch.as_tensor(rng.uniform(-1, 1, (3, B, J)).astype(np.float32)).cuda()
    res = []
    for rep in range(2):
        env.set_state(u0, 0); env.rollout_device(acts); res.append(env.get_state()[0])
    for rep in range(2):
        env.set_state(u0, 0)
        for k in range(3): env.step_device(acts[k])
        res.append(env.get_state()[0])
    env.set_state(u0, 0); env.step_device(acts[0]); a1 = env.get_state()[0]
    env.set_state(u0, 0); env.rollout_device(acts[:1]); b1 = env.get_state()[0]
    print(N, B, "roll-roll", np.abs(res[0]-res[1]).max(), "step-step", np.abs(res[2]-res[3]).max(), "roll-step", np.abs(res[0]-res[2]).max(),
          "K1", np.abs(a1-b1).max(), "rows differing", np.nonzero(np.abs(res[0]-res[2]).max(1))[0].tolist())
    env.close()
