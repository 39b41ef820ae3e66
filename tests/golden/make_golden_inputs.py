"""Generates tests/golden/config1_inputs.json from the reference's own fixtures
(/root/reference/tests/Q99536.fasta + LQSRPAAPPAPGPGQLTLR.mzML). Run in the authoring
container only (the GPU box has no /root/reference); the JSON is committed.

mzML decode rules follow crates/sage-cloudpath/src/mzml.rs:109-403: selected-ion m/z overrides
the isolation-window target (:229-231,244-248), isolation window = Da(-lower, upper) (:354-357),
f32 arrays are read little-endian after base64+zlib (:292-327), cvParam values are parsed
straight into f32 (correctly rounded from the decimal string).
"""
import base64, json, re, struct, sys, zlib
from decimal import Decimal
import numpy as np

REF = "/root/reference/tests"


def parse_f32(s: str) -> float:
    """Correctly rounded decimal -> f32 (Rust str::parse::<f32>), avoiding double rounding."""
    d = Decimal(s)
    f = np.float32(float(s))
    cands = [f, np.nextafter(f, np.float32(np.inf)), np.nextafter(f, np.float32(-np.inf))]
    best = min(cands, key=lambda c: (abs(Decimal(float(c)) - d), int(np.float32(c).view(np.uint32)) & 1))
    return float(best)


def main():
    xml = open(f"{REF}/LQSRPAAPPAPGPGQLTLR.mzML").read()
    spectra = re.findall(r"<spectrum .*?</spectrum>", xml, flags=re.S)
    assert len(spectra) == 1
    sp = spectra[0]

    def cv(acc, scope=sp):
        m = re.search(r'accession="%s"[^>]*?value="([^"]*)"' % acc, scope)
        return m.group(1) if m else None

    assert cv("MS:1000511") == "2" and 'accession="MS:1000127"' in sp  # ms level 2, centroid
    prec = re.search(r"<precursor .*?</precursor>", sp, flags=re.S).group(0)
    sel = re.search(r"<selectedIon>.*?</selectedIon>", prec, flags=re.S).group(0)
    arrays = {}
    for bda in re.findall(r"<binaryDataArray .*?</binaryDataArray>", sp, flags=re.S):
        assert "MS:1000521" in bda and "MS:1000574" in bda  # f32 + zlib
        raw = zlib.decompress(base64.b64decode(re.search(r"<binary>(.*?)</binary>", bda, flags=re.S).group(1)))
        arr = np.frombuffer(raw, dtype="<f4")
        arrays["mz" if "MS:1000514" in bda else "intensity"] = arr
    out = {
        "source": "lazear/sage @0639176 tests/Q99536.fasta + tests/LQSRPAAPPAPGPGQLTLR.mzML",
        "fasta": open(f"{REF}/Q99536.fasta").read(),
        "spectrum_id": re.search(r'<spectrum [^>]*id="([^"]*)"', sp).group(1),
        "ms_level": 2,
        "scan_start_time_min": parse_f32(cv("MS:1000016")),
        "precursor_mz": parse_f32(cv("MS:1000744", sel)),
        "precursor_charge": int(cv("MS:1000041", sel)),
        "isolation_window_da": [-parse_f32(cv("MS:1000828", prec)), parse_f32(cv("MS:1000829", prec))],
        "raw_total_ion_current": parse_f32(cv("MS:1000285")),
        # f32 values are exactly representable as Python floats; stored via their u32 bit patterns for safety
        "mz_bits": [int(x) for x in arrays["mz"].view(np.uint32)],
        "intensity_bits": [int(x) for x in arrays["intensity"].view(np.uint32)],
    }
    assert len(out["mz_bits"]) == 299 == len(out["intensity_bits"])
    json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "tests/golden/config1_inputs.json", "w"), indent=0)
    print("peaks", len(out["mz_bits"]), "precursor", out["precursor_mz"], out["precursor_charge"], out["isolation_window_da"])


if __name__ == "__main__":
    main()
