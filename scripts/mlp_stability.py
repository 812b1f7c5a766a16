import os, sys, time
for v in ("OMP_NUM_THREADS","OPENBLAS_NUM_THREADS","MKL_NUM_THREADS"): os.environ[v]="1"
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent.parent))
import numpy as np, warnings
warnings.filterwarnings("ignore")
from sklearn.neural_network import MLPClassifier
from oracle import cases, dajin2_oracle as orc

def train_set(control, threshold):
    rng = np.random.default_rng(seed=1)
    n_rand, n_ctrl = 10_000, len(control)
    base = rng.uniform(0, 100, n_rand)
    base_err = np.clip(base + rng.uniform(0, threshold, n_rand), 0, 100)
    base_mut = np.clip(base + rng.uniform(threshold, 100, n_rand), 0, 100)
    ctrl_err = np.clip(control + rng.uniform(0, threshold, n_ctrl), 0, 100)
    ctrl_mut = np.clip(control + rng.uniform(threshold, 100, n_ctrl), 0, 100)
    err = np.concatenate([np.array([base, base_err]).T, np.array([control, ctrl_err]).T], axis=0)
    mut = np.concatenate([np.array([base, base_mut]).T, np.array([control, ctrl_mut]).T], axis=0)
    X = np.concatenate([err, mut], axis=0)
    y = np.array([0] * (n_rand + n_ctrl) + [1] * (n_rand + n_ctrl))
    return X, y

def fit(X,y):
    clf = MLPClassifier(solver="lbfgs", alpha=1e-5, hidden_layer_sizes=(5, 2), random_state=1, max_iter=1000)
    t=time.perf_counter(); clf.fit(X,y); dt=time.perf_counter()-t
    return clf, dt

for name in sys.argv[1:]:
    case = cases.build_case(name)
    if "alleles" in case: case = case["alleles"]["control"]
    L=len(case["sequence"])
    ns = orc.normalize_indels(orc.count_indels(case["sample"]["rows"], L))
    nc = orc.normalize_indels(orc.count_indels(case["control"]["rows"], L))
    ncm = orc.minimize_mutation_counts(nc, ns)
    for op in "+-*":
        s=np.asarray(ns[op],float); c=np.asarray(ncm[op],float)
        X,y=train_set(c,0.1)
        clf,dt=fit(X,y)
        P=np.array([c,s]).T
        base=clf.predict(P)
        prob=clf.predict_proba(P)[:,1]
        print(name,op,"fit %.2fs n_iter %d loss %.5f"%(dt,clf.n_iter_,clf.loss_),"pos",int(base.sum()),"near(0.3-0.7)",int(((prob>0.3)&(prob<0.7)).sum()), flush=True)
        rng=np.random.default_rng(5)
        for trial in range(3):
            perm=rng.permutation(len(y))
            clf2,_=fit(X[perm],y[perm])
            p2=clf2.predict(P)
            print("   shuffle",trial,"n_iter",clf2.n_iter_,"loss %.5f"%clf2.loss_,"diff preds",int((p2!=base).sum()), flush=True)
        clf3,_=fit(X*(1+1e-13),y)
        print("   perturb diff", int((clf3.predict(P)!=base).sum()), "n_iter",clf3.n_iter_, flush=True)
