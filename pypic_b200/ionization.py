"""Monte-Carlo ionisation: rate-coefficient tables of
pygcpic.Particle.attempt_first_ionization / attempt_nth_ionization (pygcpic.py:350-458) and the
host side of the event loop.

The probabilities are computed on the device for every eligible particle
(pic_dev_gc_post_push); the uniform draws, the charge-state updates and -- because an
ionisation of a source neutral changes the running source-ion count that decides
re-activation vs deletion of the slots visited later in the same pass (pygcpic.py:1544) -- the
re-activate-or-delete decisions are made on the host, in index order, over the (short) list of
eligible and inactive slots, consuming the legacy MT19937 stream exactly like the reference.
"""
import numpy as np

# electron temperature [eV] -> rate coefficient [cm^3/s]; physical data as tabulated in the reference
_TABLES = {
    (1, 0): ([8.626E-01, 1.011E+00, 2.178E+00, 3.539E+00, 5.146E+00, 7.069E+00, 9.410E+00, 1.231E+01, 1.598E+01,
              2.076E+01, 2.720E+01, 3.625E+01, 4.973E+01, 7.133E+01, 1.099E+02, 1.904E+02, 4.079E+02, 1.355E+03,
              1.390E+04, 8.595E+04],
             [7.553E-16, 8.291E-15, 1.714E-11, 2.470E-10, 9.985E-10, 2.398E-09, 4.412E-09, 6.940E-09, 9.869E-09,
              1.309E-08, 1.649E-08, 1.996E-08, 2.329E-08, 2.624E-08, 2.834E-08, 2.881E-08, 2.627E-08, 1.926E-08,
              8.109E-09, 3.829E-09]),                                          # pygcpic.py:373-383
    (5, 0): ([8.626E-01, 1.329E+00, 2.160E+00, 3.140E+00, 4.314E+00, 5.741E+00, 7.508E+00, 9.746E+00, 1.267E+01,
              1.660E+01, 2.212E+01, 3.034E+01, 4.353E+01, 6.704E+01, 1.162E+02, 2.490E+02, 8.265E+02, 8.481E+03,
              8.669E+04],
             [1.057E-12, 3.996E-11, 5.912E-10, 2.458E-09, 6.083E-09, 1.155E-08, 1.878E-08, 2.767E-08, 3.806E-08,
              4.979E-08, 6.257E-08, 7.590E-08, 8.901E-08, 1.005E-07, 1.080E-07, 1.079E-07, 9.470E-08, 5.161E-08,
              2.159E-08]),                                                     # :409-418
    (5, 1): ([8.612E-01, 1.869E+00, 4.028E+00, 6.547E+00, 9.522E+00, 1.308E+01, 1.741E+01, 2.276E+01, 2.956E+01,
              3.840E+01, 5.031E+01, 6.707E+01, 9.203E+01, 1.319E+02, 2.033E+02, 3.522E+02, 7.547E+02, 2.505E+03,
              2.571E+04, 8.582E+04],
             [1.375E-21, 1.396E-14, 2.693E-11, 3.643E-10, 1.393E-09, 3.188E-09, 5.629E-09, 8.554E-09, 1.182E-08,
              1.533E-08, 1.900E-08, 2.273E-08, 2.639E-08, 2.972E-08, 3.221E-08, 3.300E-08, 3.032E-08, 2.252E-08,
              9.306E-09, 5.538E-09]),                                          # :420-428
    (5, 2): ([1.366E+00, 2.819E+00, 6.073E+00, 9.875E+00, 1.436E+01, 1.972E+01, 2.624E+01, 3.432E+01, 4.456E+01,
              5.790E+01, 7.587E+01, 1.012E+02, 1.387E+02, 1.990E+02, 3.064E+02, 5.311E+02, 1.138E+03, 3.778E+03,
              3.877E+04, 8.602E+04],
             [1.230E-21, 2.871E-15, 5.524E-12, 7.439E-11, 2.824E-10, 6.401E-10, 1.117E-09, 1.677E-09, 2.293E-09,
              2.946E-09, 3.629E-09, 4.337E-09, 5.055E-09, 5.759E-09, 6.382E-09, 6.779E-09, 6.575E-09, 5.269E-09,
              2.483E-09, 1.829E-09]),                                          # :429-438
}
ORDER = ((1, 0), (5, 0), (5, 1), (5, 2))


def rate(Z, charge_state, temperature_K):
    """np.interp(temperature, Te_K, R_m3_s) exactly as pygcpic.py:385-388 / 440-443."""
    Te, R = _TABLES[(int(Z), int(charge_state))]
    Te_K = [T * 11600. for T in Te]
    R_m3_s = [r / 1e6 for r in R]
    return float(np.interp(temperature_K, Te_K, R_m3_s))


def rates(temperature_K):
    return [rate(Z, c, temperature_K) for Z, c in ORDER]


def run_events(ev_idx, ev_kind, prob, cs, Zs, p2cs, midexit, A, B, source_Z, source_N, on_reactivate, rng=np.random):
    """The sequential part of one pass of the particle loop, over the event slots only.

    ev_idx (ascending), ev_kind (0 = ionisation attempt of an active particle, 1 = slot inactive at
    loop entry); prob, cs, Zs, p2cs, midexit: per event; A[e] = deterministic source-ion count of the
    slots BEFORE the event after their update, B[e] = count of the slots from the event on in their
    entry state.  Draw order = index order: one uniform per attempt (pygcpic.py:393/453; the
    ``and charge_state == 0`` test of BOTH routines is applied after the draw); on_reactivate(slot)
    is called at the right point of the stream so the caller can advance the source generator.
    Returns (ionised slots, their new charge states, p2c added by ionisations, re-activated slots,
    deleted slots)."""
    extra = 0                       # stochastic change of the running count so far in this pass
    ionised, new_cs, added, reactivated, deleted = [], [], [], [], []
    for e in range(len(ev_idx)):
        i = int(ev_idx[e])
        if ev_kind[e] == 0:
            u = rng.uniform(0., 1.)
            if u < prob[e] and cs[e] == 0.:
                ionised.append(i); new_cs.append(cs[e] + 1.); added.append(float(p2cs[e]))
                if int(Zs[e]) == source_Z and not midexit[e]:
                    extra += 1      # a source neutral became a source ion (and did not leave mid-domain)
        else:
            if A[e] + B[e] + extra < source_N:
                on_reactivate(i)
                reactivated.append(i)
                extra += 1
            else:
                deleted.append(i)
    return ionised, new_cs, added, reactivated, deleted
