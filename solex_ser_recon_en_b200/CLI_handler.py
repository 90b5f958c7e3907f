"""Command-line flags of SHG_MAIN with the reference's semantics
(/root/reference/CLI_handler.py:10-114): single-dash clusters such as
``-dcfm``, ``-w<spec>`` with spec ``a,b,c`` / ``x:y`` / ``x:y:step`` (inclusive,
attached to the flag), ``-r<width>``; unknown letters print the usage and are
skipped; non-.ser/.avi arguments are warned about and ignored."""
from __future__ import annotations

import sys

flag_dictionnary = {
    'h': 'Help',
    'w': 'shift',
    'd': 'flag_display',
    'x': 'ratio_fixe',
    'f': 'save_fit',
    'c': 'clahe_only',
    'p': 'disk_display',
    's': 'crop_width_square',
    't': 'transversalium',
    'm': 'flip_x',
    'r': 'fixed_width',
}

_SWITCHES_ON = {'d': 'flag_display', 'f': 'save_fit', 'c': 'clahe_only', 's': 'crop_width_square', 'm': 'flip_x'}


def usage():
    lines = [
        "SHG_MAIN.py [-hwdxfcpstmr] [file(s) to treat, * allowed]",
        "'h' : 'Help', display help menu.",
        "'w' : 'a,b,c, ...'  produce images at a, b, c ... pixels.",
        "'w' : 'x:y:w'  produce images starting at x, finishing at y, every w pixels.",
        "'d' : 'flag_display', display all graphics (False by default)",
        "'x' : 'ratio_fixe', disable ellipse fitting",
        "'f' : 'save_fit', save all fits files (False by default)",
        "'c' : 'clahe_only',  only final clahe image is saved (False by default)",
        "'p' : 'disk_display' turn off black disk with protuberance images (False by default)",
        "'s' : 'crop_square_width', crop the width to equal the height (False by default)",
        "'t' : 'disable transversalium', disable transversalium correction (False by default)",
        "'m' : 'mirror flip', mirror flip in x-direction (False by default)",
        "'r' : 'w'  crop width to a constant no. of pixels.",
    ]
    return '\n'.join(lines)


def _take(text, pos, allowed):
    """Longest run of characters from `allowed` starting at text[pos]."""
    end = pos
    while end < len(text) and (text[end].isdigit() or text[end] in allowed):
        end += 1
    return text[pos:end], end


def parse_shift_spec(spec):
    parts = spec.split(':')
    if len(parts) == 1:
        return [int(x.strip()) for x in spec.split(',')]
    if len(parts) == 2:
        return list(range(int(parts[0].strip()), int(parts[1].strip()) + 1))
    if len(parts) == 3:
        return list(range(int(parts[0].strip()), int(parts[1].strip()) + 1, int(parts[2].strip())))
    print('invalid shift input')
    sys.exit()


def treat_flag_at_cli(options, argument):
    """Apply one ``-xyz`` argument to the options dict (in place)."""
    options['disk_display'] = True
    body = argument[1:]
    i = 0
    while i < len(body):
        ch = body[i]
        if ch == 'h':
            print(usage())
            sys.exit()
        elif ch == 'w':
            spec, end = _take(body, i + 1, ':,-')
            i = end
            options['shift'] = parse_shift_spec(spec)
        elif ch == 'r':
            digits, end = _take(body, i + 1, '')
            i = end
            options['fixed_width'] = int(digits)
        elif ch == 't':
            options['transversalium'] = False
            i += 1
        elif ch == 'p':
            options['disk_display'] = False
            i += 1
        elif ch == 'x':
            options['ratio_fixe'] = 1
            i += 1
        elif ch in _SWITCHES_ON:
            options[_SWITCHES_ON[ch]] = True
            i += 1
        else:
            print('ERROR !!! At least one argument is not accepted')
            print(usage())
            i += 1
    print('options %s' % (options))


def handle_CLI(options, argv=None):
    argv = sys.argv[1:] if argv is None else argv
    serfiles = []
    for argument in argv:
        if argument[:1] == '-':
            treat_flag_at_cli(options, argument)
        elif argument.split('.')[-1].upper() in ('SER', 'AVI'):
            serfiles.append(argument)
        else:
            print(f'WARNING: {argument} was not a valid SER or AVI file name and was ignored. '
                  'Remember to use "-" if you want to input a flag')
    print('theses files are going to be processed : ', serfiles)
    return serfiles
