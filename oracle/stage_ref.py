"""TEST / BENCH INFRASTRUCTURE ONLY -- stage the reference's own Python modules for this path into the
git-ignored directory oracle/_ref/ so that they travel to the GPU box with the snapshot (the way the built
libshg.so does) and `bench.py --impl reference` can time the UNMODIFIED reference there.

    python -m oracle.stage_ref            # also run by __graft_entry__.build() when /root/reference is present

Nothing from the reference enters the git history: oracle/_ref/ is listed in .gitignore (but not in
.gpurunignore).  The modules are stored byte for byte in ONE archive, oracle/_ref/reference_modules.zip, which
Python imports from directly (zipimport); a MANIFEST with their sha256 is written next to it.
The modules import matplotlib / astropy / tkinter / FreeSimpleGUI / skimage / ellipse; oracle/ref_shim.py stubs
the non-numeric ones and stands in oracle/thirdparty.py for the two numeric packages that cannot be installed
here (scikit-image, lsq-ellipse -- see that file's header for how they are cross-checked).
"""
from __future__ import annotations

import hashlib
import json
import os
import zipfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCE = os.environ.get('SHG_REFERENCE_DIR', '/root/reference')
DEST = os.path.join(ROOT, 'oracle', '_ref')
MODULES = ('video_reader.py', 'solex_util.py', 'Solex_recon.py', 'ellipse_to_circle.py', 'CLI_handler.py')
ARCHIVE = os.path.join(DEST, 'reference_modules.zip')


def stage(verbose: bool = False) -> bool:
    """Archive the modules if the reference tree is present.  Returns True when oracle/_ref is usable."""
    if os.path.isfile(os.path.join(SOURCE, 'solex_util.py')):
        os.makedirs(DEST, exist_ok=True)
        blobs = {name: open(os.path.join(SOURCE, name), 'rb').read() for name in MODULES}
        manifest = {name: hashlib.sha256(b).hexdigest() for name, b in blobs.items()}
        if not verify(manifest):
            with zipfile.ZipFile(ARCHIVE, 'w', zipfile.ZIP_STORED) as z:
                for name in MODULES:
                    z.writestr(zipfile.ZipInfo(name, date_time=(2020, 1, 1, 0, 0, 0)), blobs[name])
            with open(os.path.join(DEST, 'MANIFEST.json'), 'w') as f:
                json.dump({'source': 'thelondonsmiths/Solex_ser_recon_EN (unmodified copies)', 'sha256': manifest}, f,
                          indent=1)
        if verbose:
            print('staged %d reference modules into %s' % (len(MODULES), ARCHIVE))
    return verify()


def verify(expect: dict | None = None) -> bool:
    """True when the archive holds every module with the manifest's (or `expect`'s) checksum."""
    try:
        manifest = expect or json.load(open(os.path.join(DEST, 'MANIFEST.json')))['sha256']
        with zipfile.ZipFile(ARCHIVE) as z:
            for name in MODULES:
                if hashlib.sha256(z.read(name)).hexdigest() != manifest[name]:
                    return False
        return True
    except Exception:
        return False


if __name__ == '__main__':
    print('oracle/_ref usable:', stage(verbose=True))
