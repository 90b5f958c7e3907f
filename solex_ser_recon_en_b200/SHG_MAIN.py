"""Command-line front end with the reference's contract
(/root/reference/SHG_MAIN.py:41-68 options, :98-143 precheck / handle_files,
:218-248 main):

    python -m solex_ser_recon_en_b200.SHG_MAIN [-dcfmpstx] [-w<spec>] [-r<N>] file.ser [file.avi ...]

The GUI (FreeSimpleGUI form, folder / continuous mode, spectral analyser) is
outside this package's scope: with no file arguments this prints the usage.
Under torchrun every rank runs this same command; frames are sharded across
the ranks and rank 0 writes the outputs.
"""
from __future__ import annotations

import json
import os
import sys
import traceback

from . import CLI_handler, Solex_recon, video_reader

options = {
    'language': 'English',
    'shift': [0],
    'flag_display': False,
    'ratio_fixe': None,
    'slant_fix': None,
    'save_fit': False,
    'clahe_only': False,
    'protus_only': False,
    'disk_display': True,
    'delta_radius': 0,
    'crop_width_square': False,
    'transversalium': True,
    'stubborn_transversalium': False,
    'trans_strength': 301,
    'img_rotate': 0,
    'flip_x': False,
    'workDir': '',
    'fixed_width': None,
    'output_dir': '',
    'input_dir': '',
    'specDir': '',
    'selected_mode': 'File input mode',
    'continuous_detect_mode': False,
    'dispersion': 0.05,
    'ellipse_fit_shift': 10,
    'de-vignette': False,
}


def _config_path():
    return os.path.join(os.path.dirname(sys.argv[0]), 'SHG_config.txt')


def read_ini():
    print('loading config file...')
    try:
        with open(_config_path(), 'r', encoding='utf-8') as fp:
            options.update(json.load(fp))
    except Exception:
        print('note: error reading config file - using default parameters')


def write_ini():
    if os.environ.get('SHG_NO_CONFIG'):
        return
    try:
        print('saving config file ...')
        with open(_config_path(), 'w', encoding='utf-8') as fp:
            json.dump(options, fp, sort_keys=True, indent=4)
    except Exception:
        traceback.print_exc()
        print('ERROR: failed to write config file: ' + _config_path())


def precheck_files(serfiles, options):
    options['tempo'] = 30000 if len(serfiles) == 1 else 5000
    good_tasks = []
    for serfile in serfiles:
        print(serfile)
        if serfile == '' or os.path.basename(serfile) == '':
            print('filename ERROR : ', serfile)
            continue
        try:
            open(serfile, 'rb').close()
        except Exception:
            traceback.print_exc()
            print('ERROR opening file : ', serfile)
            continue
        if not good_tasks:
            if options['selected_mode'] == 'File input mode':
                options['workDir'] = os.path.dirname(serfile) + '/'
            write_ini()
        good_tasks.append((serfile, options.copy()))
    if not good_tasks:
        write_ini()
    return good_tasks


def handle_files(files, options, flag_command_line=False):
    good_tasks = precheck_files(files, options)
    try:
        Solex_recon.solex_do_work(good_tasks, flag_command_line)
    except Exception:
        print('ERROR ENCOUNTERED')
        traceback.print_exc()
        if int(os.environ.get('WORLD_SIZE', '1')) > 1:
            # one rank per GPU: the other ranks are waiting in a collective for this one; carrying on (as the
            # single-process reference does) would leave them there until NCCL times out -- take the job down
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(1)


def is_openable(file):
    try:
        open(file, 'rb').close()
        return video_reader.video_reader(file).FrameCount > 0
    except Exception:
        return False


def _init_distributed():
    """Join the process group when launched by torchrun (one rank per GPU)."""
    if int(os.environ.get('WORLD_SIZE', '1')) > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
            dist.init_process_group('nccl')


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    serfiles = CLI_handler.handle_CLI(options, argv) if argv else []
    if not serfiles:
        print(CLI_handler.usage())
        print('(the graphical front end of the reference is not part of this package: pass SER / AVI files)')
        return 1
    _init_distributed()
    handle_files(serfiles, options, flag_command_line=True)
    return 0


if __name__ == '__main__':
    sys.exit(main())
