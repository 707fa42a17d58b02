"""slimm_b200.collect_profiles against outputs of the reference's collect_profiles.py (run in the build container on profiles that
are themselves outputs of the reference binary: tests/golden/*/runs/*/profile.tsv; the merged files are committed under
tests/golden/merged/)."""
import os
import shutil

from slimm_b200 import collect_profiles

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SETS = {
    "toyA_toyB_quirk": [("toy_yara/runs/default", "toyA_profile.tsv"), ("toy_yara/runs/w1000", "toyB_profile.tsv"),
                        ("quirk/runs/default", "quirk_profile.tsv")],
    # a dot inside a file name and a directory with a dot: the sample name is cut at the last '.' of the whole path
    "lca64genus_lca64species_synth1k": [("lca64/runs/cc1_genus", "g_profile.tsv"), ("lca64/runs/cc1_species", "s.x_profile.tsv"),
                                        ("synth1k/runs/w1000", "sub.dir/k_profile.tsv")],
}


def _stage(tmp_path, files):
    out = []
    for src, name in files:
        dst = tmp_path / name
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copy(os.path.join(GOLD, src, "profile.tsv"), dst)
        out.append(name)
    return out


def test_merge_is_byte_identical_to_the_reference_script(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    for name, files in SETS.items():
        paths = _stage(tmp_path, files)
        assert collect_profiles.main(paths) == 0                  # writes merged_profile.tsv into the working directory, like the reference
        exp = open(os.path.join(GOLD, "merged", name + ".merged_profile.tsv"), "rb").read()
        assert open(tmp_path / "merged_profile.tsv", "rb").read() == exp, name


def test_clean_merge(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    paths = _stage(tmp_path, SETS["toyA_toyB_quirk"][:2])
    assert collect_profiles.main(["--clean", "-o", "clean.tsv"] + paths) == 0
    rows = [ln.split("\t") for ln in open(tmp_path / "clean.tsv").read().splitlines()]
    assert rows[0] == ["taxa_level", "taxa_id", "linage", "toyA_profile.read_count", "toyA_profile.abundance",
                       "toyB_profile.read_count", "toyB_profile.abundance"]
    by_id = {r[1]: r for r in rows[1:]}
    assert by_id["131"][3:5] == ["4740", "23.7059"] and by_id["0*"][3] == "2637"
    assert sum(int(r[3]) for r in rows[1:]) == 19995            # every read of sample A is in exactly one row
