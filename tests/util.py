"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np

from oracle import oracle as O


class DeviceState:
    """velm / posq / force of a DrudeSystem as CUDA tensors in the boundary layouts."""

    def __init__(self, system, device, force_format=0, padded=None):
        import torch
        n = system.num_particles
        self.n = n
        self.padded = padded or ((n + 31) // 32) * 32
        self.torch = torch
        velm = np.zeros((self.padded, 4), np.float32); velm[:n] = system.velm_f32()
        posq = np.zeros((self.padded, 4), np.float32); posq[:n] = system.posq_f32(charges=np.arange(n) % 7 - 3.0)
        self.velm = torch.from_numpy(velm).to(device)
        self.posq = torch.from_numpy(posq).to(device)
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
        return self.posq[: self.n, :3].double().cpu().numpy()

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
