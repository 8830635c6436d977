"""
Recipe for oracle/_ref/ -- TEST INFRASTRUCTURE ONLY (git-ignored, never part of the product path).

The reference's implementation of the hot path is one pure-Python file,
/root/reference/tangram/mapping_optimizer.py (imports numpy, logging, torch only: lines 9-12).  There is
nothing to compile; to let the UNMODIFIED reference run where /root/reference does not exist (the GPU box),
this script copies that one file verbatim into oracle/_ref/ next to a stamp with its sha256.  oracle/_ref/ is
listed in .gitignore (the reference's source never enters this repo's history) but not in .gpurunignore, so
the copy travels with the snapshot like a built .so does.

Users (and only these): bench.py's reference arms (`--impl reference` on the host cores; the `reference_gpu`
comparator with device='cuda') and the parity tests that run the live reference beside the CUDA path.

    python oracle/build_ref.py            # no-op when /root/reference is absent (uses the existing copy)
"""
import hashlib
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/tangram/mapping_optimizer.py"
REF_DIR = os.path.join(HERE, "_ref")
REF_DST = os.path.join(REF_DIR, "mapping_optimizer.py")
STAMP = os.path.join(REF_DIR, "SOURCE.txt")


def build():
    """Copy the reference file if the reference tree is present.  Returns the path of the copy or None."""
    if os.path.exists(REF_SRC):
        os.makedirs(REF_DIR, exist_ok=True)
        shutil.copyfile(REF_SRC, REF_DST)
        with open(REF_DST, "rb") as f:
            digest = hashlib.sha256(f.read()).hexdigest()
        with open(STAMP, "w") as f:
            f.write(f"verbatim copy of {REF_SRC}\nsha256 {digest}\n")
    return REF_DST if os.path.exists(REF_DST) else None


def load():
    """The unmodified reference module (classes Mapper, MapperConstrained), loaded by path."""
    import importlib.util
    path = REF_DST if os.path.exists(REF_DST) else (REF_SRC if os.path.exists(REF_SRC) else None)
    if path is None:
        raise FileNotFoundError("oracle/_ref/mapping_optimizer.py is missing: run `python oracle/build_ref.py` "
                                "where /root/reference exists (it travels to the GPU box with the snapshot)")
    spec = importlib.util.spec_from_file_location("tangram_reference_mapping_optimizer", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build() or "no reference tree and no existing copy")
