"""Synthetic GNSS epochs for the per-epoch preprocessing tests (tests/test_gnss_epoch.py): one receiver near
(22.3 N, 114.2 E), 20 satellites of three systems with integer ambiguities, receiver clocks, cycle slips (announced by
the slip counter), unannounced jumps (caught by the median gate), a satellite below the elevation mask, an unhealthy
one, and a satellite that disappears for longer than the ambiguity timeout.  Measurements follow the residual models
of RVI/factor/gnss_factor.cpp so that every residual is zero at the truth up to the added noise."""
import ctypes as C

import numpy as np

import swgn_gnss as G

OMGE, CLIGHT = 7.2921151467E-5, 299792458.0
RE, FE = 6378137.0, 1.0 / 298.257223563


def pos2ecef(lat, lon, h):
    e2 = FE * (2.0 - FE)
    v = RE / np.sqrt(1.0 - e2 * np.sin(lat) ** 2)
    return np.array([(v + h) * np.cos(lat) * np.cos(lon), (v + h) * np.cos(lat) * np.sin(lon), (v * (1.0 - e2) + h) * np.sin(lat)])


def enu_basis(lat, lon):
    return np.array([[-np.sin(lon), np.cos(lon), 0.0],
                     [-np.sin(lat) * np.cos(lon), -np.sin(lat) * np.sin(lon), np.cos(lat)],
                     [np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)]])


def sagnac_range(rr, rs):
    return np.linalg.norm(rr - rs) + OMGE * (rs[0] * rr[1] - rs[1] * rr[0]) / CLIGHT


def range_rate(rr, rs, vr, vs):
    e = (rr - rs) / np.linalg.norm(rr - rs)
    return (vr - vs) @ e + OMGE / CLIGHT * (vs[1] * rr[0] + rs[1] * vr[0] - vs[0] * rr[1] - rs[0] * vr[1])


class Scenario:
    def __init__(self, seed=0, n_epochs=14, cfg=None, pose_noise=0.01, unhealthy_has_phase=True, half_flag=1):
        rng = np.random.default_rng(seed)
        self.rng = rng
        self.cfg = cfg
        self.lams = np.array([[cfg.lams[s][f] for f in range(2)] for s in range(3)])
        lat, lon = np.radians(22.3), np.radians(114.17)
        self.base = np.round(pos2ecef(lat, lon, 40.0), 3)
        E = enu_basis(lat, lon)
        self.n_epochs = n_epochs
        self.pose_noise = pose_noise
        # the reference asserts an ambiguity for every non-zero RTK phase once rover_count > 1 (swf_core.cpp:110), unhealthy
        # satellites included: real data carries no phase for them
        self.unhealthy_has_phase = unhealthy_has_phase
        self.half_flag = half_flag   # LambdaSearch asserts bits 8 and 2 of it (swf_lambda.cpp:161)
        # satellites: system, number, ENU direction, range, velocity
        sys_of = [0] * 8 + [1] * 7 + [2] * 5
        first = [1, 40, 77]
        self.sats = []
        for k, s in enumerate(sys_of):
            az = rng.uniform(0, 2 * np.pi)
            el = np.radians(rng.uniform(32, 85))
            if k == 3:
                el = np.radians(12.0)   # below AZELMIN for the whole run
            d = np.array([np.sin(az) * np.cos(el), np.cos(az) * np.cos(el), np.sin(el)]) @ E
            pos = self.base + rng.uniform(2.0e7, 2.6e7) * d
            vel = np.cross(d, rng.normal(size=3))
            vel *= 3000.0 / np.linalg.norm(vel)
            self.sats.append(dict(sys=s, sat=first[s] + sum(1 for x in sys_of[:k] if x == s), pos=pos, vel=vel,
                                  N=float(rng.integers(-1000, 1000)), N_spp=float(rng.integers(-1000, 1000)), slip=0, spp_slip=0,
                                  svh=1 if k == 10 else 0))
        self.p0 = np.array([12.0, -7.0, 3.0])
        self.v0 = np.array([1.5, 0.7, -0.1])
        self.clk = rng.uniform(-30, 30, 13)
        self.clk_rate = rng.uniform(-0.5, 0.5, 13)
        self.clk[12] = rng.uniform(-2, 2)
        self.clk_rate[12] = 0.0
        # events: (epoch, satellite index, kind)
        self.events = {(4, 1): "slip", (6, 9): "jump", (8, 16): "slip", (3, 5): "spp_jump"}
        self.gone = {12: (5, 3)}  # satellite 12 is missing for epochs 5..7 (shorter than the ambiguity timeout)

    def truth_pose(self, t):
        return self.p0 + self.v0 * t

    def epoch(self, k):
        """(Epoch, obs array keep-alive, Frame) of epoch k; the frame carries the true pose plus pose_noise."""
        rng = self.rng
        t = float(k) * 1.0 + (12.0 if k >= 9 else 0.0)  # a 13 s gap before epoch 9: every ambiguity times out
        p = self.truth_pose(t)
        rr = p + self.base
        obs = (G.Obs * len(self.sats))()
        n = 0
        for i, s in enumerate(self.sats):
            ev = self.events.get((k, i))
            if ev == "slip":
                s["slip"] += 1
                s["N"] += float(rng.integers(5, 50))
            if ev == "jump":
                s["N"] += 3.0   # unannounced: the slip counter stays
            if ev == "spp_jump":
                s["N_spp"] += 4.0
            if i in self.gone and self.gone[i][0] <= k < self.gone[i][0] + self.gone[i][1]:
                continue
            lam = self.lams[s["sys"]][0]
            pos = s["pos"] + s["vel"] * t
            rho = sagnac_range(rr, pos)
            o = obs[n]
            n += 1
            o.sat, o.sys, o.svh = s["sat"], s["sys"], s["svh"]
            o.rtk_slip_count[0] = s["slip"] & 255
            o.spp_slip_count[0] = s["spp_slip"] & 255
            o.half_flag[0] = self.half_flag
            clk_rtk, clk_spp = self.clk[s["sys"] * 2] + self.clk_rate[s["sys"] * 2] * t, self.clk[6 + s["sys"] * 2] + self.clk_rate[6 + s["sys"] * 2] * t
            o.rtk_l[0] = (rho - s["N"] * lam + clk_rtk) / lam + rng.normal(0, 0.003)
            o.rtk_p[0] = rho + clk_rtk + rng.normal(0, 0.25)
            o.rtk_lstd[0], o.rtk_pstd[0] = 0.01, 0.3
            o.spp_p[0] = rho + clk_spp + rng.normal(0, 0.8)
            o.spp_l[0] = (rho - s["N_spp"] * lam + clk_spp) / lam + rng.normal(0, 0.01)
            o.spp_pstd[0], o.spp_lstd[0] = 0.6, 0.02
            if s["svh"] and not self.unhealthy_has_phase:
                o.rtk_l[0] = o.spp_l[0] = 0.0
            rate = range_rate(rr, pos, self.v0, s["vel"])
            o.spp_d[0] = -(rate + self.clk[12]) / lam + rng.normal(0, 0.02)
            o.spp_dstd[0] = 0.05
            for c in range(3):
                o.sat_pos[c], o.sat_vel[c] = pos[c], s["vel"][c]
            o.el = 0.0
            o.sat_var, o.ion_var, o.trop_var = 0.5, 1.5, 0.3
        e = G.Epoch()
        e.n_obs, e.ros_time, e.br_time_diff = n, 1000.0 + t, 0.4
        for c in range(3):
            e.base_xyz[c] = self.base[c]
        e.obs = C.cast(obs, C.POINTER(G.Obs))
        f = G.Frame()
        pn = p + rng.normal(0, self.pose_noise, 3)
        q = np.array([0.1, -0.2, 0.3, 0.9])
        q /= np.linalg.norm(q)
        for c in range(3):
            f.pose[c] = pn[c]
            f.speed_bias[c] = self.v0[c] + rng.normal(0, 0.01)
        for c in range(4):
            f.pose[3 + c] = q[c]
        for c in range(6):
            f.speed_bias[3 + c] = 0.01 * (c + 1)
        f.blackvalue = 0.02
        f.nonlinear, f.rover_count, f.epochs_since_start, f.not_fix_count = 1, min(k + 1, 10), k, 0
        return e, obs, f


def copy_epoch(e, obs):
    """Deep copy (the preprocessing modifies epochs in place): (Epoch, obs keep-alive)."""
    o2 = (G.Obs * len(obs))()
    C.memmove(o2, obs, C.sizeof(obs))
    e2 = G.Epoch()
    C.memmove(C.byref(e2), C.byref(e), C.sizeof(G.Epoch))
    e2.obs = C.cast(o2, C.POINTER(G.Obs))
    return e2, o2


def copy_frame(f):
    f2 = G.Frame()
    C.memmove(C.byref(f2), C.byref(f), C.sizeof(G.Frame))
    return f2
