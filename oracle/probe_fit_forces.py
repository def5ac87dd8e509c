"""Study aid (test infrastructure): fit the estimator's toe / heel force model.  Model under test (consistent with the archive to 1e-7 in a local
test): toeForce = heelForce = -1/2 J_c^-T [k_s shin; k_h H], k = (1500, 1250), H(knee, shin, tarsus) = heel-spring angle that closes the achilles-rod
four-bar, J_c = Jacobian of the foot point (pelvis x-z) w.r.t. the two spring deflections under that closure.  Unknown: the four-bar geometry."""
import sys
import numpy as np
from scipy.optimize import brentq, least_squares
from probe_est_tools import est, fk, Rz, frame

D13 = np.deg2rad(13)
HS_POS = np.array([-0.01269, -0.03059, 0.00092])
HS_FRAME = frame((-0.91211, 0.40829, 0.036948), (-0.40992, -0.90952, -0.068841))


def orth(F):
    u, _, vt = np.linalg.svd(F)
    return u @ vt


HS_FRAME = orth(HS_FRAME)


def rod_gap(par, kn, sh, ta, dh):
    """|A - B| - L in the hip-pitch frame; planar chain knee -> shin -> tarsus -> heel spring (angle dh about its own z)"""
    A, Bl, Lr = par[0:3], par[3:6], par[6]
    p = np.array([0.12, 0, 0.0045]); R = Rz(kn)
    p = p + R @ np.array([0.06068, 0.04741, 0]); R = R @ Rz(sh)
    p = p + R @ np.array([0.43476, 0.02, 0]); R = R @ Rz(ta)
    p = p + R @ HS_POS; R = R @ HS_FRAME @ Rz(dh)
    B = p + R @ Bl
    return np.linalg.norm(A - B) - Lr


def H(par, kn, sh, ta):
    return brentq(lambda d: rod_gap(par, kn, sh, ta, d), -0.6, 0.6, xtol=1e-15)


def force_model(par, kn, sh, ta, ft=-1.5):
    h = 1e-6
    Hv = H(par, kn, sh, ta)
    a = (H(par, kn, sh, ta + h) - H(par, kn, sh, ta - h)) / (2 * h)
    b = (H(par, kn, sh + h, ta) - H(par, kn, sh - h, ta)) / (2 * h)
    m = [0, 0, 0, kn, ft]
    ps = ((fk(0, m, sh + h, ta)[0] - fk(0, m, sh - h, ta)[0]) / (2 * h))[[0, 2]]
    pt = ((fk(0, m, sh, ta + h)[0] - fk(0, m, sh, ta - h)[0]) / (2 * h))[[0, 2]]
    Jc = np.stack([ps - pt * b / a, pt / a], axis=1)
    tau = np.array([1500.0 * sh, 1250.0 * Hv])
    return -0.5 * np.linalg.solve(Jc.T, tau)


def measured(kn, sh, ta, ft=-1.5):
    x = np.zeros(45); x[32] = 1; x[3] = kn; x[4] = ft; x[20] = sh; x[21] = ta; x[8] = -1.2; x[9] = -1.5; x[24] = D13 + 1.2
    return est(x)[35:38][[0, 2]]


if __name__ == '__main__':
    rng = np.random.default_rng(0)
    cfgs = [(rng.uniform(-2.3, -0.8), rng.uniform(-0.04, 0.04), None, rng.uniform(-0.04, 0.04)) for _ in range(40)]
    cfgs = [(kn, sh, D13 - kn - sh + dh) for kn, sh, _, dh in cfgs]
    meas = np.array([measured(*c) for c in cfgs])
    par0 = np.array([0, 0, 0.045, 0.11943353, -0.00866334, -0.00156846, 0.5012])

    def resid(par):
        try:
            return np.concatenate([force_model(par, *c) - m for c, m in zip(cfgs, meas)])
        except ValueError:
            return np.full(2 * len(cfgs), 1e3)
    r0 = resid(par0)
    print('MJCF geometry: max |F model - F archive| = %.3e (force scale %.1f)' % (np.abs(r0).max(), np.abs(meas).max()))
    sol = least_squares(resid, par0, x_scale=[0.01, 0.01, 0.01, 0.01, 0.01, 0.01, 0.01], xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=200)
    print('fitted', np.round(sol.x, 7), 'max resid %.3e' % np.abs(sol.fun).max())
