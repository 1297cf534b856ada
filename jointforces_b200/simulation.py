"""``jf.simulation`` call surface on top of the CUDA relaxation library.

Keeps the signatures, file names and file formats of ``/root/reference/jointforces/simulation.py``
for the solver path (``read_meshfile`` :25, ``spherical_contraction_solver`` :81,
``distribute_solver`` :171, ``extract_deformation_curve_solver`` :263,
``create_lookup_table_solver`` :306, ``create_lookup_functions`` :595, ``save_lookup_table`` :639,
``load_lookup_functions`` :643, ``linear_lookup_interpolator`` :703) so that the reference's own
table builder and ``jf.force.reconstruct`` consume the output folders unchanged.

What differs is only *how* a sweep runs: the reference forks one OS process per pressure, each
re-parsing the mesh and rebuilding the solver (``simulation.py:218-246``); here the mesh is
parsed and analysed once per GPU and the pressures assigned to a GPU are relaxed in lockstep
batches by the same kernels (see ``sweep.py``).
"""
from __future__ import annotations

import os
from glob import glob

import numpy as np

from . import _lib
from .solver import Solver, SemiAffineFiberMaterial, load_solver  # noqa: F401  (re-exported like the reference)


# --------------------------------------------------------------------------------------------------
# mesh file (reference simulation.py:25-78)
# --------------------------------------------------------------------------------------------------
def _parse_block(lines, start, count, width, after_first=False):
    """Rows ``start .. start+count`` of ``lines`` as a ``(count, width)`` float array of their last ``width`` tokens
    (elements, reference :70) or of all tokens after the first (nodes, reference :60 -- then there must be exactly
    ``width`` of them).  Vectorised when every row has the same token count, else the reference's line loop."""
    if count == 0:
        return np.zeros((0, width))
    first = lines[start].split()
    try:
        flat = np.array(" ".join(lines[start:start + count]).split(), dtype=np.float64)
        if flat.size != count * len(first) or (after_first and len(first) != width + 1) or len(first) < width:
            raise ValueError
        return flat.reshape(count, len(first))[:, -width:]
    except ValueError:
        out = np.zeros((count, width))
        for i in range(count):
            tokens = lines[start + i].split()
            out[i] = np.array([float(x) for x in tokens[1:]]) if after_first else tokens[-width:]
        return out


def read_meshfile(meshfile, r_inner=None, r_outer=None):
    """Parse a Gmsh 2.2 ASCII mesh.  Returns ``coords (N,3) [m], tets (T,4) float 0-based,
    r_inner [m], r_outer [m]`` exactly like the reference: radii found in a ``$Jointforces``
    block override the arguments, every ``$Elements`` line contributes its last four tokens."""
    fast = _read_meshfile_fast(meshfile, r_inner, r_outer)
    if fast is not None:
        return fast
    return _read_meshfile_lines(meshfile, r_inner, r_outer)


def _read_meshfile_fast(meshfile, r_inner, r_outer):
    """The regular case through the library's threaded reader (jf_msh_*, bit-identical numbers); None when the
    library leaves the file to the line loop (anything but plain decimal ASCII) or has not been built."""
    try:
        res = _lib.read_msh(meshfile)
    except _lib.JFError:
        res = None
    if res is None:
        return None
    coords, tets, radii = res
    if radii is not None:
        r_inner, r_outer = radii
        print('Geometry for spherical contraction is defined in mesh file (r_inner={:.2f}, r_outer={:.2f}).'.format(r_inner, r_outer))
        print('Will ignore user-defined values of r_inner and r_outer.')
    else:
        if r_inner is None:
            raise ValueError('r_inner not defined')
        if r_outer is None:
            raise ValueError('r_outer not defined')
    return coords, tets, r_inner * 10 ** -6, r_outer * 10 ** -6


def _read_meshfile_lines(meshfile, r_inner=None, r_outer=None):
    """The reference's line loop (simulation.py:25-78), vectorised per block where the rows are regular."""
    with open(meshfile, 'r') as f:
        lines = f.readlines()

    try:
        a = lines.index('$Jointforces\n')
        b = lines.index('$EndJointforces\n')
        meshinfo = {}
        for line in lines[a + 1:b]:
            key, value = line.strip().split('=')
            try:
                meshinfo[key] = float(value)
            except ValueError:
                meshinfo[key] = value
        r_inner = meshinfo['r_inner']
        r_outer = meshinfo['r_outer']
        print('Geometry for spherical contraction is defined in mesh file (r_inner={:.2f}, r_outer={:.2f}).'.format(r_inner, r_outer))
        # the reference prints this whenever the block exists (its test runs after the overwrite, :41-46)
        print('Will ignore user-defined values of r_inner and r_outer.')
    except ValueError:
        if r_inner is None:
            raise ValueError('r_inner not defined')
        if r_outer is None:
            raise ValueError('r_outer not defined')

    r_inner = r_inner * 10 ** -6
    r_outer = r_outer * 10 ** -6

    i_nodes = lines.index('$Nodes\n')
    n_nodes = int(lines[i_nodes + 1])
    coords = np.ascontiguousarray(_parse_block(lines, i_nodes + 2, n_nodes, 3, after_first=True))

    i_elem = lines.index('$Elements\n')
    n_elem = int(lines[i_elem + 1])
    tets = np.ascontiguousarray(_parse_block(lines, i_elem + 2, n_elem, 4))
    tets -= 1
    return coords, tets, r_inner, r_outer


# --------------------------------------------------------------------------------------------------
# boundary conditions and parameter file of one solve (reference simulation.py:107-164)
# --------------------------------------------------------------------------------------------------
def spherical_boundary_conditions(coords, r_inner, r_outer, pressure):
    """NaN-coded conditions of a contracting inclusion: outer shell fixed at zero displacement,
    every inner-shell node loaded with the same force ``pressure * 4 pi r_inner^2 / N_inner``
    along its radius (uniform per node, not area weighted -- as the reference)."""
    distance = np.sqrt(np.sum(coords ** 2., axis=1))
    mask_inner = distance < r_inner * 1.001
    mask_outer = distance > r_outer * 0.999

    bcond_displacement = np.full((len(coords), 3), np.nan)
    bcond_displacement[mask_outer] = 0

    bcond_forces = np.zeros((len(coords), 3))
    bcond_forces[mask_outer] = np.nan
    n_inner = np.sum(mask_inner)
    force_per_node = pressure * (4 * np.pi * r_inner ** 2.) / n_inner
    bcond_forces[mask_inner] = coords[mask_inner] / distance[mask_inner, None] * force_per_node

    spacing_inner = np.sqrt((np.pi * 4 * r_inner ** 2) / n_inner)
    spacing_outer = np.sqrt((np.pi * 4 * r_outer ** 2) / np.sum(mask_outer))
    return dict(displacement=bcond_displacement, forces=bcond_forces, mask_inner=mask_inner, mask_outer=mask_outer,
                force_per_node=force_per_node, inner_spacing=spacing_inner, outer_spacing=spacing_outer)


_PARAMETER_KEYS = ("K_0", "D_0", "L_S", "D_S", "PRESSURE", "FORCE_PER_SURFACE_NODE", "INNER_RADIUS", "OUTER_RADIUS",
                   "INNER_NODE_SPACING", "OUTER_NODE_SPACING", "SURFACE_NODES", "TOTAL_NODES")


def write_parameters(outfolder, material, pressure, bc, r_inner, r_outer, n_nodes):
    """``parameters.txt``: twelve ``KEY = value`` lines, lengths carrying a `` µm`` suffix
    (reference simulation.py:149-164; parsed back at :265-275)."""
    values = (material['K_0'], material['D_0'], material['L_S'], material['D_S'], pressure, bc['force_per_node'],
              r_inner * 1e6, r_outer * 1e6, bc['inner_spacing'] * 1e6, bc['outer_spacing'] * 1e6,
              np.sum(bc['mask_inner']), n_nodes)
    rows = []
    for key, value in zip(_PARAMETER_KEYS, values):
        unit = " µm" if key.endswith("RADIUS") or key.endswith("SPACING") else ""
        rows.append("{} = {}{}".format(key, value, unit))
    with open(outfolder + "/parameters.txt", "w") as f:
        f.write("\n".join(rows))


def spherical_contraction_solver(meshfile, outfolder, pressure, material, r_inner=None, r_outer=None, logfile=False,
                                 initial_displacenemts=None, max_iter=600, step=0.0033, conv_crit=0.01):
    """One relaxation of a contracting spherical inclusion; writes ``parameters.txt``,
    ``relrec.txt`` and ``solver.saenopy`` into ``outfolder`` (reference simulation.py:81-168; the
    misspelt ``initial_displacenemts`` is the reference's keyword)."""
    coords, tets, r_inner, r_outer = read_meshfile(meshfile, r_inner, r_outer)

    if not os.path.exists(outfolder):
        os.makedirs(outfolder)
    else:
        print('WARNING: Output folder already exists! ({})'.format(outfolder))

    M = Solver()
    M.set_material_model(SemiAffineFiberMaterial(material['K_0'], material['D_0'], material['L_S'], material['D_S']))
    M.set_nodes(coords)
    M.set_tetrahedra(tets)

    bc = spherical_boundary_conditions(coords, r_inner, r_outer, pressure)
    print('Inner node spacing: ' + str(bc['inner_spacing'] * 1e6) + 'µm')
    print('Outer node spacing: ' + str(bc['outer_spacing'] * 1e6) + 'µm')
    M.set_boundary_condition(bc['displacement'], bc['forces'])
    if initial_displacenemts is not None:
        M.set_initial_displacements(initial_displacenemts)

    write_parameters(outfolder, material, pressure, bc, r_inner, r_outer, len(coords))

    M.solve_boundarycondition(step_size=step, max_iterations=max_iter, rel_conv_crit=conv_crit,
                              relrecname=outfolder + "/relrec.txt")
    M.save(outfolder + "/solver.saenopy")
    # the legacy archive the reference's last loader fallback reads with any saenopy version (:252-258, 289-291)
    save_legacy = getattr(M, "save_legacy", None)   # (an object with saenopy's own Solver protocol has no such method)
    if save_legacy is not None:
        save_legacy(outfolder + "/solver.npz")
    return M


def distribute_solver(func, const_args, var_arg='pressure', start=0.1, end=1000, n=120, log_scaling=True, n_cores=None,
                      get_initial=True, max_iter=600, step=0.0033, conv_crit=0.01, callback=None):
    """Pressure sweep with the reference's signature and output layout (simulation.py:171-249):
    ``<outfolder>/pressure-values.txt`` and one ``simulation%06d/`` folder per value.

    ``func`` is accepted and ignored like the reference does (:175).  ``n_cores`` bounds how many
    pressures are relaxed in lockstep on one GPU (default: library maximum).  ``get_initial`` is
    accepted; solves are cold starts, which is what the reference does in effect (it looks for a
    ``solver.npz`` that the current saenopy never writes, :168 vs :210,229).
    Under ``torchrun`` the pressures are sharded over the ranks (one GPU each)."""
    from .sweep import run_sweep

    if "max_iter" in const_args:
        max_iter = const_args["max_iter"]
    if "step" in const_args:
        step = const_args["step"]
    if "conv_crit" in const_args:
        conv_crit = const_args["conv_crit"]

    if log_scaling:
        values = np.logspace(np.log10(start), np.log10(end), n, endpoint=True)
    else:
        # the reference's linear branch spaces the *logarithms* linearly (:192); kept
        values = np.linspace(np.log10(start), np.log10(end), n, endpoint=True)

    outfolder = const_args['outfolder']
    del const_args['outfolder']  # the reference mutates the caller's dict (:195); kept

    run_sweep(const_args["meshfile"], outfolder, values, const_args["material"], var_arg=var_arg, max_iter=max_iter,
              step=step, conv_crit=conv_crit, batch=n_cores, callback=callback,
              r_inner=const_args.get("r_inner"), r_outer=const_args.get("r_outer"))
    return


# --------------------------------------------------------------------------------------------------
# lookup table (reference simulation.py:263-323, 595-659, 703-736)
# --------------------------------------------------------------------------------------------------
def radial_median_curve(coords, displ, r_inner, x):
    """Median of |u|/r_inner over the nodes whose |x|/r_inner falls in each bin [x_i, x_i+1)
    (reference simulation.py:296-299); empty bins give NaN."""
    u = np.sqrt(np.sum(coords ** 2., axis=1)) / r_inner
    v = np.sqrt(np.sum(displ ** 2., axis=1)) / r_inner
    order = np.argsort(u, kind="stable")
    u_sorted, v_sorted = u[order], v[order]
    edges = np.searchsorted(u_sorted, x, side="left")
    y = np.full(len(x) - 1, np.nan)
    for i in range(len(x) - 1):
        if edges[i + 1] > edges[i]:
            y[i] = np.nanmedian(v_sorted[edges[i]:edges[i + 1]])
    return y


def read_parameters(folder):
    parameters = {}
    with open(folder + '/parameters.txt', 'r') as f:
        for line in f.readlines():
            try:
                key, value = line.split('= ')
                parameters[key.strip()] = float(value.split(' ')[0].strip())
            except Exception:
                pass
    return parameters


def extract_deformation_curve_solver(folder, x):
    parameters = read_parameters(folder)
    try:
        M = load_solver(folder + "/solver.saenopy")
    except Exception:
        M = load_solver(folder + "/solver.npz")
    y = radial_median_curve(M.mesh.nodes, M.mesh.displacements, parameters['INNER_RADIUS'] * 10 ** -6, x)
    return {'pressure': parameters['PRESSURE'], 'bins': x, 'displacements': y}


def lookup_bins(x0=1, x1=50, n=100):
    x = np.logspace(np.log10(x0), np.log10(x1), n + 1, endpoint=True)
    x_center = 10 ** (0.5 * (np.log10(x[1:]) + np.log10(x[:-1])))
    return x, x_center


def create_lookup_table_solver(folder, x0=1, x1=50, n=100):
    """``{'pressure': (n_sims,), 'x': (n,), 'y': (n_sims, n)}`` from the ``simulation*/`` folders
    of a sweep (reference simulation.py:306-323).  Folders are read in sorted order (the
    reference's ``glob`` order is arbitrary; the interpolants do not depend on row order)."""
    subfolders = sorted(glob(folder + '/*/'))
    x, x_center = lookup_bins(x0, x1, n)
    pressure_values, displacement_curves = [], []
    for subfolder in subfolders:
        res = extract_deformation_curve_solver(subfolder, x)
        pressure_values.append(res['pressure'])
        displacement_curves.append(res['displacements'])
    return {'pressure': np.array(pressure_values), 'x': x_center, 'y': np.array(displacement_curves)}


def create_lookup_functions(lookup_table):
    from scipy.interpolate import LinearNDInterpolator
    pressure = np.asarray(lookup_table['pressure'])
    distance = np.asarray(lookup_table['x'])
    displacement = np.asarray(lookup_table['y'])

    gx, gy = np.meshgrid(np.log(distance), np.log(pressure))
    ok = ~(np.isnan(gx) | np.isnan(gy) | np.isnan(displacement))
    gx, gy, displacement = gx[ok], gy[ok], displacement[ok]

    forward = LinearNDInterpolator(np.array([gx, gy]).T, displacement)
    inverse = LinearNDInterpolator(np.array([gx, displacement]).T, gy)

    def get_displacement(distance, pressure):
        return forward(np.log(distance), np.log(pressure))

    def get_pressure(distance, displacement):
        return np.exp(inverse(np.log(distance), displacement))

    return get_displacement, get_pressure


def save_lookup_table(lookup_table, outfile):
    np.save(outfile, lookup_table)


def load_lookup_functions(file):
    lookup_table = np.load(file, allow_pickle=True).item()
    get_displacement, get_pressure = create_lookup_functions(lookup_table)
    print("Found Lookuptable, and created lookup functions")
    return get_displacement, get_pressure


def linear_lookup_interpolator(emodulus, output_newtable="new-lin-lookup.npy", reference_folder=None):
    """Rescale the shipped K_0 = 2364 linear table to Young's modulus ``emodulus`` using
    K_0 = 6 E (reference simulation.py:703-736).  ``reference_folder`` must hold
    ``linear_reference_k2364.npy`` (it ships with the reference's docs/data, not with this repo)."""
    if not reference_folder:
        raise ValueError("reference_folder with linear_reference_k2364.npy is required")
    lookup_table = np.load(os.path.join(reference_folder, 'linear_reference_k2364.npy'), allow_pickle=True).item()
    shifted = dict(lookup_table)
    shifted['y'] = np.array(lookup_table['y']) * 2364 / (emodulus * 6)
    save_lookup_table(shifted, output_newtable)
    return shifted
