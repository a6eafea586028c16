"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np

from oracle import oracle as O


class DeviceState:
    """velm / posq / force of a DrudeSystem as CUDA tensors in the boundary layouts.
    precision 0: OpenMM single (float4 velm, float4 posq); 1: OpenMM mixed (double4 velm, float4 posq + float4 posqCorrection);
    2: OpenMM double (double4 velm, double4 posq)."""

    def __init__(self, system, device, force_format=0, padded=None, precision=0):
        import torch
        n = system.num_particles
        self.n = n
        self.padded = padded or ((n + 31) // 32) * 32
        self.torch = torch
        self.precision = precision
        vt = np.float64 if precision else np.float32
        velm = np.zeros((self.padded, 4), vt)
        velm[:n, :3] = system.velocities
        velm[:n, 3] = system.inv_masses.astype(np.float32) if not precision else system.inv_masses
        posq = np.zeros((self.padded, 4), np.float64 if precision == 2 else np.float32)
        posq[:n, :3] = system.positions
        posq[:n, 3] = np.arange(n) % 7 - 3.0
        self.velm = torch.from_numpy(velm).to(device)
        self.posq = torch.from_numpy(posq).to(device)
        self.corr = None
        if precision == 1:
            corr = np.zeros((self.padded, 4), np.float32)
            corr[:n, :3] = system.positions - posq[:n, :3].astype(np.float64)
            self.corr = torch.from_numpy(corr).to(device)
        self.force_format = force_format
        self.set_forces(system.forces)
        self.charges = posq[:n, 3].copy()

    def set_forces(self, forces):
        torch = self.torch
        n = self.n
        if self.force_format == 0:
            f = np.zeros((3, self.padded), np.float32); f[:, :n] = forces.T
        else:
            f = np.zeros((3, self.padded), np.int64); f[:, :n] = np.rint(forces.T * 4294967296.0).astype(np.int64)
        self.force = torch.from_numpy(f).to(self.velm.device)

    def vel(self):
        return self.velm[: self.n, :3].double().cpu().numpy()

    def pos(self):
        p = self.posq[: self.n, :3].double()
        if self.corr is not None:
            p = p + self.corr[: self.n, :3].double()
        return p.cpu().numpy()

    @property
    def ptrs(self):
        return self.velm.data_ptr(), self.posq.data_ptr(), self.force.data_ptr()


def rel_err(a, ref):
    """max |a - ref| / max(|ref|, rms(ref)): relative error that does not blow up on near-zero components."""
    scale = np.maximum(np.abs(ref), np.sqrt(np.mean(ref * ref)) + 1e-300)
    return float(np.max(np.abs(a - ref) / scale))


def group_temperatures(ke2, dof):
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(dof > 0, ke2 / (dof * O.BOLTZ), 0.0)


def ke_err(got, ref, nkbt):
    """Error of the per-thermostat 2*KE sums in units that matter to the thermostat: relative to the group's own
    energy scale max(|2KE_g|, N_g kT).  A group without degrees of freedom (N_g kT = 0, e.g. ions whose only
    relative motion is the Drude pair itself) has 2KE = 0 exactly and no thermostat; there the error is taken
    relative to the total so that fp32 cancellation noise (1e-9 of the total) is not divided by zero."""
    scale = np.maximum(np.abs(ref), np.abs(nkbt))
    scale = np.where(scale > 1e-3 * np.abs(ref).max(), scale, np.abs(ref).max())
    return float(np.max(np.abs(got - ref) / scale))


def chain_err(got, ref):
    """Chain variables relative to the largest magnitude in the array: eta_dot_0 of a thermostat in equilibrium is a
    small difference (KE - N kT) / Q of large numbers, so its own magnitude is not a meaningful scale."""
    return float(np.max(np.abs(got - ref)) / max(np.abs(ref).max(), 1e-300))
