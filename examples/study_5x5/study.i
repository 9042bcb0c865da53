# 5 x 5 x 1 HEX8 study on [-5, 5]^3 with a different material in the centre element: the shape of
# MaCaw's own test problem (test/tests/study/openmc_study.i), written for the synthetic nuclide
# tables of this repository.  Run with
#   host/mcw_run --input examples/study_5x5/study.i --tables lib.bin      (lib.bin: python -m macaw_b200.problem_io --library lib.bin)
[Mesh]
  [gmg]
    type = GeneratedMeshGenerator
    dim = 3
    nx = 5
    ny = 5
    nz = 1
    xmin = -5
    ymin = -5
    zmin = -5
    xmax = 5
    ymax = 5
    zmax = 5
  []
  [add_subdomain]
    input = gmg
    type = SubdomainBoundingBoxGenerator
    top_right = '1 1 5'
    bottom_left = '-1 -1 -5'
    block_id = 1
    block_name = 'center'
  []
[]

[Problem]
  solve = false
  kernel_coverage_check = false
[]

[Variables/temperature]
  order = CONSTANT
  family = MONOMIAL
  initial_condition = 293.6
[]

[AuxVariables/flux]
  order = CONSTANT
  family = MONOMIAL
[]

[RayKernels/collision]
  type = CollisionKernel
  temperature = temperature
  blocks = "0 1"
  materials = "4 1"  # OpenMC material ids: borated water around a fuel element
[]

[RayBCs]
  [reflect]
    type = ReflectRayBC
    boundary = 'back front top right left bottom'
  []
[]

[UserObjects/study]
  type = OpenMCStudy
  execute_on = TIMESTEP_END
[]

[UserObjects/flux_tally]
  type = OpenMCTally
  id = 7
  estimator = COLLISION
  scores = 'flux nu-fission'
  filters = 'cell energy'
  energy_bins = '1e-5 0.625 2e7'
  execute_on = INITIAL
[]

[AuxKernels/thermal_flux]
  type = OpenMCTallyAux
  variable = flux
  tally_id = 7
  granularity = cell
  score = flux
  estimator = COLLISION
  particle_type = neutron
  energy_bin = 0
[]

[Executioner]
  type = Transient
  num_steps = 4
[]

[Outputs]
  csv = true
[]
