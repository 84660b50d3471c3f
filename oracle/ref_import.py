"""Import the *real* reference (samrere/pytortto, `tortto` v1.3.4) numpy path in this container.

TEST INFRASTRUCTURE ONLY.  Used by `oracle/make_golden.py` to generate the committed fixtures under
`tests/golden/` and by `tests/test_oracle_vs_reference.py` (skipped when /root/reference is absent,
i.e. on the GPU box).  Nothing in `pytortto_b200/` may import this module.

Recipe:
  * copy /root/reference/src/tortto to a writable temp dir (the package regenerates grad_fcn.py on import,
    reference `autograd/grad_fcn_generator.py:34-35`),
  * patch `xparray.py:14` (numpy 2 removed `np._get_promotion_state`; NEP-50 "weak" is the default there),
  * patch `autograd/grad_nn.py:789,819` (`xp.NINF` was removed in numpy 2 -> `-xp.inf`),
  * import it BEFORE torch/scipy (reference `xparray.py:10-12` deletes and re-imports numpy).
No reference source is copied into this repository; the patched copy lives in a temp dir only.
"""
import os
import shutil
import sys
import tempfile

REFERENCE_SRC = os.environ.get("TORTTO_REFERENCE_SRC", "/root/reference/src/tortto")


def reference_available():
    return os.path.isdir(REFERENCE_SRC)


def import_reference():
    """Returns the imported, patched reference package `tortto` (numpy path)."""
    if "tortto" in sys.modules:
        return sys.modules["tortto"]
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_SRC}")
    for m in ("torch", "scipy"):
        if m in sys.modules:
            raise RuntimeError(f"import the reference before {m} (reference xparray.py:10-12 re-imports numpy)")
    tmp = tempfile.mkdtemp(prefix="tortto_ref_")
    dst = os.path.join(tmp, "tortto")
    shutil.copytree(REFERENCE_SRC, dst)
    for root, _dirs, files in os.walk(dst):
        os.chmod(root, 0o755)
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)
    p = os.path.join(dst, "xparray.py")
    s = open(p).read()
    s = s.replace("assert np._get_promotion_state() == 'weak', \"numpy import error\"",
                  "assert getattr(np, '_get_promotion_state', lambda: 'weak')() == 'weak', \"numpy import error\"")
    open(p, "w").write(s)
    p = os.path.join(dst, "autograd", "grad_nn.py")
    s = open(p).read()
    s = s.replace("xp.NINF", "(-xp.inf)")
    open(p, "w").write(s)
    sys.path.insert(0, tmp)
    import tortto  # noqa: E402
    return tortto
