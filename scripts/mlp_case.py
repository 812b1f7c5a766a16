import numpy as np
def train_set(control, thr):
    rng = np.random.default_rng(seed=1)
    n_rand, n_ctrl = 10_000, len(control)
    base = rng.uniform(0, 100, n_rand)
    base_err = np.clip(base + rng.uniform(0, thr, n_rand), 0, 100)
    base_mut = np.clip(base + rng.uniform(thr, 100, n_rand), 0, 100)
    ctrl_err = np.clip(control + rng.uniform(0, thr, n_ctrl), 0, 100)
    ctrl_mut = np.clip(control + rng.uniform(thr, 100, n_ctrl), 0, 100)
    err = np.concatenate([np.array([base, base_err]).T, np.array([control, ctrl_err]).T], axis=0)
    mut = np.concatenate([np.array([base, base_mut]).T, np.array([control, ctrl_mut]).T], axis=0)
    X = np.concatenate([err, mut], axis=0)
    y = [0] * (n_rand + n_ctrl) + [1] * (n_rand + n_ctrl)
    return X, y
def case(L=5000, seed=5):
    rng0=np.random.default_rng(seed)
    control=np.abs(rng0.normal(0.8,0.2,L)); sample=control+rng0.normal(0,0.05,L); sample[[L//5,L//3,L//2]]+= [25,15,10]
    control=np.minimum(control,np.clip(sample,0,100))
    return train_set(control,0.1)
