"""Generates tests/golden/config1_fullrun.npz (build container only: reads /root/reference/GalaxyCollision.IC and runs the
UNMODIFIED reference program oracle/_ref/Gadget2_ref_f32 built by integration/Makefile):
  in_vel, masstab          the velocity block and mass table of the shipped example (positions/types are in config1_galaxycollision.npz)
  ref_*                    what the pure-CPU reference program produces after TimeMax = 0.05 (32 steps): every 16th particle's
                           position and velocity, the energy.txt rows, the number of steps and force computations."""
import os
import struct
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "integration"))


def main():
    f = open("/root/reference/GalaxyCollision.IC", "rb")

    def block():
        n = struct.unpack("i", f.read(4))[0]
        d = f.read(n)
        assert struct.unpack("i", f.read(4))[0] == n
        return d
    h = block()
    masstab = np.asarray(struct.unpack("6d", h[24:72]))
    block()
    vel = np.frombuffer(block(), dtype=np.float32).reshape(-1, 3).copy()
    out = os.path.join(HERE, "config1_fullrun.npz")
    np.savez_compressed(out, in_vel=vel, masstab=masstab)
    import fullrun
    with tempfile.TemporaryDirectory() as tmp:
        r = fullrun.run("ref", "f32", tmp, time_max=0.05)
    s = r["snap"]
    np.savez_compressed(out, in_vel=vel, masstab=masstab, ref_time_max=0.05, ref_stride=16, ref_pos=s["pos"][::16], ref_vel=s["vel"][::16],
                        ref_energy=r["energy"], ref_steps=r["steps"], ref_force_computations=r["force_computations"],
                        ref_ia_per_part=r["ia_per_part_mean"])
    print(out, os.path.getsize(out), "steps", r["steps"], "wall", r["wall_s"], "gravity", r["cpu_gravity_s"])


if __name__ == "__main__":
    main()
