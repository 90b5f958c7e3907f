"""Wall-clock of the drop-in CLI on BASELINE configs[0..3] at full size (files on tmpfs): the same
measurement bench.py prints as `config_lines` (see bench.config_lines), runnable on its own.

    python tools/config_bench.py [max_files_for_config_4]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                              # noqa: E402
from solex_ser_recon_en_b200.engine import get_engine                    # noqa: E402

print(json.dumps(bench.config_lines(get_engine(0), int(sys.argv[1]) if len(sys.argv) > 1 else 16), indent=1))
