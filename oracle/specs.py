"""TEST INFRASTRUCTURE ONLY -- the model specs behind tests/golden/*.npz
(shared by oracle/make_golden.py and tests/)."""
import math

PI = math.pi

# name -> Spec kwargs.  Covers: deep lattice / free / ideal / defects,
# N even / odd / not a multiple of 4, N != L, non-integer L, near and far
# Jastrow branches.
SPECS = {
    'deep_n100': dict(lattice_depth=100, lattice_ratio=1,
                      interaction_strength=1, boson_number=100,
                      supercell_size=100, tbf_contact_cutoff=25),
    'll_n16': dict(lattice_depth=0, lattice_ratio=1, interaction_strength=4,
                   boson_number=16, supercell_size=16, tbf_contact_cutoff=4),
    'lat_n50': dict(lattice_depth=5 * PI ** 2, lattice_ratio=1,
                    interaction_strength=2, boson_number=50,
                    supercell_size=50, tbf_contact_cutoff=12.5),
    'lat_n100': dict(lattice_depth=5 * PI ** 2, lattice_ratio=1,
                     interaction_strength=2, boson_number=100,
                     supercell_size=100, tbf_contact_cutoff=25),
    'deep_n200': dict(lattice_depth=20 * PI ** 2, lattice_ratio=1,
                      interaction_strength=2, boson_number=200,
                      supercell_size=200, tbf_contact_cutoff=50),
    'defects_n20': dict(lattice_depth=5 * PI ** 2, lattice_ratio=0.5,
                        interaction_strength=3, boson_number=20,
                        supercell_size=20, tbf_contact_cutoff=5,
                        num_defects=4, defect_magnitude=2 * PI ** 2),
    'ideal_n8': dict(lattice_depth=5 * PI ** 2, lattice_ratio=1,
                     interaction_strength=0, boson_number=8,
                     supercell_size=8, tbf_contact_cutoff=2),
    'odd_n7': dict(lattice_depth=3 * PI ** 2, lattice_ratio=2.0,
                   interaction_strength=1.5, boson_number=7,
                   supercell_size=10, tbf_contact_cutoff=3.3),
    'frac_n21': dict(lattice_depth=2 * PI ** 2, lattice_ratio=0.25,
                     interaction_strength=8, boson_number=21,
                     supercell_size=17.5, tbf_contact_cutoff=1.75),
    'strong_n10': dict(lattice_depth=0, lattice_ratio=1,
                       interaction_strength=200, boson_number=10,
                       supercell_size=10, tbf_contact_cutoff=4.5),
}
