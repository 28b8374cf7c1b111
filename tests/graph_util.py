"""Start-up ordering of a process graph: a SINK only waits for SOURCEs that have touched its node (Sink.h:93-116),
so the frame server must not start before every consumer has.  The reference's example scripts sleep; the tests ask."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oat_b200", "bin")


def wait_sources(*addrs, n=1, timeout_s=60):
    """Block until each address has at least n SOURCEs attached (shmemdf_test wait-sources)."""
    r = subprocess.run([os.path.join(BIN, "shmemdf_test"), "wait-sources", str(n), str(timeout_s)] + list(addrs))
    assert r.returncode == 0, f"no SOURCE attached to {addrs} within {timeout_s} s"
