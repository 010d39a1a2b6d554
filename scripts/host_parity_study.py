"""Host-side parity study (CPU only): how far does a change of arithmetic in the QP stage move the converged control
points, compared with how far the reference moves against itself from x0 + 1 ulp?  Compares the single-lane host build of
the kernel source (tests/hostsim) and a VARIANT host build given as a .so (built from a patched copy of csrc/ with the
command of tests/hostsim_loader.py) with the CPU reference on the first problems of a synthetic batch.
usage: host_parity_study.py CONFIG SAMPLE [variant.so]"""
import sys, numpy as np, types, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p_)
import hostsim_loader, bench


def main():
    from trajectory_generator_b200 import synthetic as syn
    name = sys.argv[1]; sample = int(sys.argv[2])
    base = hostsim_loader.load(); var = hostsim_loader.HostSim(sys.argv[3] if len(sys.argv) > 3 else '/tmp/dot2_hostsim.so')
    bt = syn.make(name, 4096); L = bt.layout
    pool = bench.CpuPool(name, bt.take(range(sample)), os.cpu_count() or 1)
    dt, res = pool.run()
    dt2, res2 = pool.run(perturb=1)
    pool.close()
    print("reference solves: %.1f s + %.1f s" % (dt, dt2))
    st_ref = np.array([r[1] for r in res]); x_ref = np.array([r[4] for r in res])
    st2 = np.array([r[1] for r in res2]); x2 = np.array([r[4] for r in res2])
    k = L.ia + 1
    stable = (st_ref == 0) & (st2 == 0) & (np.abs(x2[:, :k] - x_ref[:, :k]).max(1) <= 1e-5)
    def run(hs):
        X=[]; S=[]
        for i in range(sample):
            pp = types.SimpleNamespace(spec=bt.spec, par=np.ascontiguousarray(bt.par[i]), x0=bt.x0[i], layout=L, xl=np.full(L.n,-np.inf), xu=np.full(L.n,np.inf))
            r = hs.solve(pp, fd=True); X.append(r['x']); S.append(r['status'])
        return np.array(X), np.array(S)
    xb, sb = run(base); xv, sv = run(var)
    def report(label, xg, sg):
        d = np.abs(xg[:, :k] - x_ref[:, :k]).max(1)
        both = (st_ref == 0) & (sg == 0)
        print("%-34s same status %d/%d | both converged %d, within 1e-5: %d | reference-stable %d, within 1e-5: %d" % (
            label, (sg == st_ref).sum(), sample, both.sum(), (d[both] <= 1e-5).sum(), stable.sum(), (d[stable & (sg == 0)] <= 1e-5).sum()))
    report("host build (kernel source)", xb, sb)
    report("host build, 2-accumulator products", xv, sv)
    dv = np.abs(xv[:, :k] - xb[:, :k]).max(1); bothc = (sb == 0) & (sv == 0)
    print("variant vs base: identical %d, within 1e-8 %d, within 1e-5 %d of %d both-converged; status differs on %d" % (
        (dv == 0).sum(), (dv[bothc] <= 1e-8).sum(), (dv[bothc] <= 1e-5).sum(), bothc.sum(), (sb != sv).sum()))


if __name__ == '__main__':
    main()
