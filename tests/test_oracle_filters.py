"""The reference's filter expectations against the CPU oracle (CPU only)."""
import pytest

import filter_cases


@pytest.mark.parametrize("case", filter_cases.ALL, ids=lambda f: f.__name__)
def test_oracle_filters(oracle, case):
    class F:
        IIRFilter = oracle.IIRFilter
        FIRFilter = oracle.FIRFilter
        FilterDesign = oracle.FilterDesign

        class FilterFactory:
            createIIRLowpass = staticmethod(lambda fc, fs: oracle.IIRFilter(**oracle.FilterDesign.butterworthLowpass(fc, fs)))
            createIIRHighpass = staticmethod(lambda fc, fs: oracle.IIRFilter(**oracle.FilterDesign.butterworthHighpass(fc, fs)))
            createIIRBandpass = staticmethod(lambda f0, bw, fs: oracle.IIRFilter(**oracle.FilterDesign.butterworthBandpass(f0, bw, fs)))
            createFIRLowpass = staticmethod(lambda fc, fs, n=51: oracle.FIRFilter(oracle.FilterDesign.sincLowpass(fc, fs, n)))
            createFIRHighpass = staticmethod(lambda fc, fs, n=51: oracle.FIRFilter(oracle.FilterDesign.sincHighpass(fc, fs, n)))
            createFIRBandpass = staticmethod(lambda f0, bw, fs, n=51: oracle.FIRFilter(oracle.FilterDesign.sincBandpass(f0, bw, fs, n)))

    case(F)


def test_design_constants_from_survey(oracle):
    """coefficient values of the FSKCore filter set."""
    lp = oracle.FilterDesign.butterworthLowpass(300, 48000)
    assert lp["b"][0] == pytest.approx(3.7506961629696616e-4, rel=1e-13)
    assert lp["a"][1] == pytest.approx(-1.9444776577670937, rel=1e-14)
    assert lp["a"][2] == pytest.approx(0.9459779362322814, rel=1e-14)
    lp = oracle.FilterDesign.butterworthLowpass(1200, 48000)
    assert lp["b"][0] == pytest.approx(5.542717210280682e-3, rel=1e-13)
    bp = oracle.FilterDesign.butterworthBandpass(1750, 2600, 48000)
    assert bp["b"][0] == pytest.approx(0.1430310558532532, rel=1e-13) and bp["b"][1] == 0 and bp["b"][2] == -bp["b"][0]
    assert bp["a"][1] == pytest.approx(-1.6691646533202396, rel=1e-14)
    # structure the CUDA kernels rely on (exact identities of the designs)
    assert lp["b"][1] == 2 * lp["b"][0] and lp["b"][2] == lp["b"][0]
