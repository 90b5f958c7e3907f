"""CPU: the command-line parser against the reference's own parser (imported
from /root/reference when present, else against recorded expectations)."""
import contextlib
import io

import pytest

from oracle import ref_shim
from solex_ser_recon_en_b200 import CLI_handler as C

CASES = [['-dcfm'], ['-w-10:10:1'], ['-w1,2,3'], ['-w-5:5'], ['-w3'], ['-tpx'], ['-r1200'], ['-sr800'], ['-w1,2c'],
         ['-w1:3:1m'], ['-cw2,4f'], ['-q'], ['-w'], ['-r'], ['-mw-2:2:2t'], ['-h'], ['-r12s'], ['-w-10:10:1', '-ms'],
         ['-w-50:50:1'], ['-w1:2:3:4']]


def run(mod, args):
    o = ref_shim.default_options(clahe_only=False)
    o.pop('_nolog')
    for a in args:
        with contextlib.redirect_stdout(io.StringIO()):
            try:
                mod.treat_flag_at_cli(o, a)
            except SystemExit:
                o['_exit'] = True
            except Exception as e:           # the reference raises ValueError on an empty spec
                o['_exc'] = type(e).__name__
    return o


@pytest.mark.skipif(not ref_shim.available(), reason='reference tree not present')
@pytest.mark.parametrize('args', CASES)
def test_flags_match_reference(args):
    assert run(C, args) == run(ref_shim.load().CLI_handler, args)


def test_known_flag_semantics():
    o = run(C, ['-w-10:10:1', '-ms'])
    assert o['shift'] == list(range(-10, 11)) and o['flip_x'] and o['crop_width_square']
    o = run(C, ['-w-50:50:1'])
    assert len(o['shift']) == 101
    o = run(C, ['-tpxr900'])
    assert o['transversalium'] is False and o['disk_display'] is False and o['ratio_fixe'] == 1 and o['fixed_width'] == 900
    assert run(C, ['-w'])['_exc'] == 'ValueError'        # "-w -10:10:1" with a space fails upstream too


def test_file_arguments(capsys):
    o = ref_shim.default_options()
    files = C.handle_CLI(o, ['-c', 'a.SER', 'b.avi', 'notes.txt', 'c.ser'])
    assert files == ['a.SER', 'b.avi', 'c.ser']
    assert 'notes.txt was not a valid SER or AVI' in capsys.readouterr().out


def test_device_image_unknown_attribute_fails_without_touching_the_pixels():
    """A name ndarray does not have must raise AttributeError at once: going through the host copy first (as
    every ndarray attribute legitimately does) would cost a full device -> host copy of the image."""
    from solex_ser_recon_en_b200.device_image import DeviceImage

    class _NoCopy(DeviceImage):
        def numpy(self):
            raise AssertionError('device -> host copy for an attribute probe')

    img = _NoCopy.__new__(_NoCopy)
    assert getattr(img, 'not_an_ndarray_attribute', None) is None
    assert not hasattr(img, 'fitfuture')
    try:
        img.astype                                            # a real ndarray attribute does go to the host copy
    except AssertionError:
        pass
    else:
        raise AssertionError('astype should have asked for the host copy')
