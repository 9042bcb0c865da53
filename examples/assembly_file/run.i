# 17 x 17 assembly on an Exodus file mesh, the shape of examples/beavrs_assembly/run_assembly.i in
# MaCaw: FileMeshGenerator + TransformGenerator, CollisionKernel on raw subdomain ids (some of which
# the mesh does not contain), reflecting side sets 1-4.  assembly.e is written from
# tests/golden/beavrs_assembly_mesh.npz (the reference's own mesh) with macaw_b200.exodus.write_exodus:
#   python -c "from macaw_b200 import exodus, meshgen; exodus.write_exodus('examples/assembly_file/assembly.e', meshgen.beavrs_assembly_file('tests/golden/beavrs_assembly_mesh.npz'))"
#   python -m macaw_b200.problem_io --library lib.bin
#   host/mcw_run --input examples/assembly_file/run.i --tables lib.bin
[Mesh]
  [core]
    type = FileMeshGenerator
    file = 'assembly.e'
  []
  [shift_z]
    type = TransformGenerator
    input = 'core'
    transform = 'TRANSLATE'
    vector_value = '0 0 100'
  []
[]

[Problem]
  solve = false
  kernel_coverage_check = false
[]

[AuxVariables]
  [temperature]
    order = CONSTANT
    family = MONOMIAL
    initial_condition = 300
  []
  [power]
    order = CONSTANT
    family = MONOMIAL
  []
[]

[RayKernels/collision]
  type = CollisionKernel
  temperature = temperature
  blocks = "10000 10004 10006 10010 10014 11006"
  materials = "3 2 3 1 4 3"  # OpenMC material ids
  z_coord = 100
[]

[RayBCs]
  [reflect]
    type = ReflectRayBC
    boundary = '1 2 3 4'
  []
[]

[UserObjects]
  [study]
    type = OpenMCStudy
    execute_on = TIMESTEP_END
  []
  [celltally]
    type = OpenMCTally
    id = 1
    particle_type = 'neutron'
    estimator = 'COLLISION'
    scores = 'kappa-fission'
    filters = 'cell'
    execute_on = 'initial'
  []
[]

[AuxKernels]
  [cell_val]
    type = OpenMCTallyAux
    granularity = 'cell'
    score = 'kappa-fission'
    tally_id = 1
    estimator = COLLISION
    particle_type = neutron
    execute_on = TIMESTEP_END
    variable = power
  []
[]

[Executioner]
  type = Transient
[]

[Outputs]
  csv = true
[]
