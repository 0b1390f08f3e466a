"""ipfitting.jl_b200 -- B200-native LsqDB assembly + lsqfit solve behind the
IPFitting.jl API.  Host-side mirror of the reference's Julia
interface; the compute lives in csrc/ (CUDA, sm_100a) behind include/ipf_b200.h."""
from .prototypes import (Atoms, Dat, ENERGY, FORCES, VIRIAL, configtype, observation,
                         hasobservation, energy, forces, virial)
from .datatypes import vec_obs, devec_obs, weighthook, err_weighthook
from .obsiter import observations
from .basis import (CosCut, BondAngleDesc, NBPoly, NBodyBasis, nbpolys, bodyorder, degree,
                    as_basis, rnn, invariant_orbits, species_nbpolys, envpolys)
from .potentials import OneBody, NBodyIP, SumIP
from .lsq_db import LsqDB, matrows, filter_basis, filter_configs, dbpath
from .lsq import lsqfit, get_lsq_system, collect_observations
from .lsqerrors import add_fits, rmse, mae
from . import synth, parallel, data
from .data import read_xyz, write_xyz, load_data
from .preconditioners import AdjacancyPrecon
