// examples/dambreak_ck.cpp — the 3-D dam break written against the C++ host layer, statement for statement in the
// order of the reference case file tests/tests_sycl/3d_examples/test_3d_dambreak_sycl/dambreak.cpp:69-225.
// Build (no nvcc needed; the GPU is reached through the C ABI of libsphb200.so only):
//   g++ -O2 -std=c++17 -Iinclude examples/dambreak_ck.cpp -Lsphinxsys_b200 -lsphb200 -Wl,-rpath,$PWD/sphinxsys_b200 -o dambreak_ck
// Run: ./dambreak_ck [dp=0.05] [end_time=1.0] [output folder for the .vtp body states; default: none written]
#include <chrono>
#include <iomanip>

#include "sphinxsys_ck/sphinxsys_ck.h"
using namespace SPH;

// geometry and material parameters (dambreak.cpp:11-25)
static Real DL = 5.366, DH = 2.0, DW = 0.5, LL = 2.0, LH = 1.0, LW = 0.5;
static Real global_resolution = 0.05;
static Real rho0_f = 1.0, gravity_g = 1.0;

class WaterBlock : public ComplexShape
{
  public:
    explicit WaterBlock(const std::string &shape_name) : ComplexShape(shape_name)
    {
        Vecd halfsize_water(0.5 * LL, 0.5 * LH, 0.5 * LW);
        add<GeometricShapeBox>(Transform(halfsize_water), halfsize_water);
    }
};
class WallBoundary : public ComplexShape
{
  public:
    explicit WallBoundary(const std::string &shape_name) : ComplexShape(shape_name)
    {
        Real BW = global_resolution * 4;
        Vecd halfsize_inner(0.5 * DL, 0.5 * DH, 0.5 * DW);
        Vecd halfsize_outer(0.5 * DL + BW, 0.5 * DH + BW, 0.5 * DW + BW);
        add<GeometricShapeBox>(Transform(halfsize_inner), halfsize_outer);
        subtract<GeometricShapeBox>(Transform(halfsize_inner), halfsize_inner);
    }
};

// observation points (dambreak.cpp:54-65)
static StdVec<Vecd> createObservationPoints()
{
    StdVec<Vecd> observation_points;
    observation_points.push_back(Vecd(DL, 0.01, 0.5 * DW));
    observation_points.push_back(Vecd(DL, 0.1, 0.5 * DW));
    observation_points.push_back(Vecd(DL, 0.2, 0.5 * DW));
    observation_points.push_back(Vecd(DL, 0.24, 0.5 * DW));
    observation_points.push_back(Vecd(DL, 0.252, 0.5 * DW));
    observation_points.push_back(Vecd(DL, 0.266, 0.5 * DW));
    return observation_points;
}

int main(int ac, char *av[])
{
    if (ac > 1) global_resolution = Real(std::atof(av[1]));
    Real end_time = ac > 2 ? Real(std::atof(av[2])) : Real(1.0);
    const bool write_states = ac > 3;
    const std::string output_folder = ac > 3 ? av[3] : "./output";
    Real BW = global_resolution * 4;
    Real U_f = 2.0 * std::sqrt(gravity_g * LH), c_f = 10.0 * U_f;
    //	Build up an SPHSystem.
    BoundingBoxd system_domain_bounds(Vecd(-BW, -BW, -BW), Vecd(DL + BW, DH + BW, DW + BW));
    SPHSystem sph_system(system_domain_bounds, global_resolution);
    //	Creating bodies with corresponding materials and particles.
    WaterBlock initial_water_block("WaterBody");
    FluidBody water_block(sph_system, initial_water_block);
    water_block.defineMatterMaterial<WeaklyCompressibleFluid>(rho0_f, c_f);
    water_block.generateParticles<BaseParticles, Lattice>();

    SolidBody wall_boundary(sph_system, makeShared<WallBoundary>("WallBoundary"));
    wall_boundary.defineMatterMaterial<Solid>();
    wall_boundary.generateParticles<BaseParticles, Lattice>();

    ObserverBody fluid_observer(sph_system, "FluidObserver");
    fluid_observer.generateParticles<ObserverParticles>(createObservationPoints());
    //	Define body relation map.
    Inner<> water_block_inner(water_block);
    Contact<> water_wall_contact(water_block, {&wall_boundary});
    Contact<> fluid_observer_contact(fluid_observer, {&water_block});
    //	Define the numerical methods used in the simulation.
    UpdateCellLinkedList<MainExecutionPolicy, RealBody> water_cell_linked_list(water_block);
    UpdateCellLinkedList<MainExecutionPolicy, RealBody> wall_cell_linked_list(wall_boundary);
    UpdateRelation<MainExecutionPolicy, Inner<>, Contact<>> water_block_update_complex_relation(water_block_inner, water_wall_contact);
    UpdateRelation<MainExecutionPolicy, Contact<>> fluid_observer_contact_relation(fluid_observer_contact);
    ParticleSortCK<MainExecutionPolicy> particle_sort(water_block);

    Gravity gravity(Vec3d(0.0, -gravity_g, 0.0));
    StateDynamics<MainExecutionPolicy, GravityForceCK<Gravity>> constant_gravity(water_block, gravity);
    StateDynamics<MainExecutionPolicy, fluid_dynamics::AdvectionStepSetup> water_advection_step_setup(water_block);
    StateDynamics<MainExecutionPolicy, fluid_dynamics::UpdateParticlePosition> water_update_particle_position(water_block);

    InteractionDynamicsCK<MainExecutionPolicy, LinearCorrectionMatrixComplex>
        fluid_linear_correction_matrix(DynamicsArgs(water_block_inner, 0.5), water_wall_contact);
    InteractionDynamicsCK<MainExecutionPolicy, fluid_dynamics::AcousticStep1stHalfWithWallRiemannCorrectionCK>
        fluid_acoustic_step_1st_half(water_block_inner, water_wall_contact);
    InteractionDynamicsCK<MainExecutionPolicy, fluid_dynamics::AcousticStep2ndHalfWithWallRiemannCorrectionCK>
        fluid_acoustic_step_2nd_half(water_block_inner, water_wall_contact);
    InteractionDynamicsCK<MainExecutionPolicy, fluid_dynamics::CompressionSummation<Inner<>, Contact<>>>
        fluid_density_summation(water_block_inner, water_wall_contact);
    StateDynamics<MainExecutionPolicy, fluid_dynamics::DensityRegularization<SPHBody, WeaklyCompressibleFluid, FreeSurface>>
        fluid_density_regularization(water_block);
    InteractionDynamicsCK<MainExecutionPolicy, fluid_dynamics::FreeSurfaceIndicationComplexSpatialTemporalCK>
        fluid_boundary_indicator(water_block_inner, water_wall_contact);
    ReduceDynamicsCK<MainExecutionPolicy, fluid_dynamics::AdvectionTimeStepCK> fluid_advection_time_step(water_block, U_f);
    ReduceDynamicsCK<MainExecutionPolicy, fluid_dynamics::AcousticTimeStepCK<WeaklyCompressibleFluid>> fluid_acoustic_time_step(water_block);
    ReduceDynamicsCK<MainExecutionPolicy, TotalMechanicalEnergyCK> record_water_mechanical_energy(water_block, gravity);
    // NormalFromBodyShapeCK (dambreak.cpp:119,153): a host dynamics in the reference too; it registers NormalDirection,
    // which the recording below puts on its write list
    wall_boundary.computeNormalFromBodyShape();
    //	Define the methods for I/O operations, observations and regression tests (dambreak.cpp:141-145).
    BodyStatesRecordingToVtpCK<MainExecutionPolicy> body_states_recording(sph_system, output_folder);
    body_states_recording.setStateRecording(write_states);
    body_states_recording.addToWrite<Vecd>(wall_boundary, "NormalDirection");
    body_states_recording.addToWrite<Real>(water_block, "Density");
    body_states_recording.addToWrite<int>(water_block, "Indicator");
    body_states_recording.addToWrite<Real>(water_block, "PositionDivergence");
    ObservedQuantityRecording<MainExecutionPolicy, Real> fluid_observer_pressure(fluid_observer_contact, "Pressure");
    //	Prepare the simulation with cell linked list, configuration and case specified initial condition.
    SingleVariable<Real> *sv_physical_time = sph_system.getSystemVariableByName<Real>("PhysicalTime");
    constant_gravity.exec();

    water_cell_linked_list.exec();
    wall_cell_linked_list.exec();
    water_block_update_complex_relation.exec();
    fluid_observer_contact_relation.exec();
    //	Setup for time-stepping control
    size_t number_of_iterations = 0, acoustic_steps = 0;
    int screen_output_interval = 100;
    Real output_interval = end_time / 20.0;
    auto t1 = std::chrono::steady_clock::now();
    std::cout << "N_fluid = " << water_block.TotalRealParticles() << "  N_wall = " << wall_boundary.TotalRealParticles()
              << "  E0 = " << std::setprecision(9) << record_water_mechanical_energy.exec() << "\n";
    body_states_recording.writeToFile(); // first output before the main loop, dambreak.cpp:177
    fluid_observer_pressure.writeToFile(number_of_iterations);
    //	Main loop starts here.
    while (sv_physical_time->getValue() < end_time)
    {
        Real integration_time = 0.0;
        while (integration_time < output_interval)
        {
            fluid_density_summation.exec();
            fluid_density_regularization.exec();
            water_advection_step_setup.exec();
            Real advection_dt = fluid_advection_time_step.exec();
            fluid_boundary_indicator.exec();
            fluid_linear_correction_matrix.exec();

            Real relaxation_time = 0.0;
            Real acoustic_dt = 0.0;
            while (relaxation_time < advection_dt)
            {
                acoustic_dt = fluid_acoustic_time_step.exec();
                fluid_acoustic_step_1st_half.exec(acoustic_dt);
                fluid_acoustic_step_2nd_half.exec(acoustic_dt);
                relaxation_time += acoustic_dt;
                integration_time += acoustic_dt;
                sv_physical_time->incrementValue(acoustic_dt);
                ++acoustic_steps;
            }
            water_update_particle_position.exec();

            if (number_of_iterations % screen_output_interval == 0)
            {
                std::cout << std::fixed << std::setprecision(9) << "N=" << number_of_iterations << "	Time = "
                          << sv_physical_time->getValue() << "	advection_dt = " << advection_dt
                          << "	acoustic_dt = " << acoustic_dt << "\n";
            }
            number_of_iterations++;

            if (number_of_iterations % 100 == 0 && number_of_iterations != 1)
            {
                particle_sort.exec();
            }
            water_cell_linked_list.exec();
            water_block_update_complex_relation.exec();
            fluid_observer_contact_relation.exec();
            fluid_observer_pressure.writeToFile(number_of_iterations);
        }
        body_states_recording.writeToFile(); // dambreak.cpp:230: device -> host synchronisation of the write list, then the file
        std::cout << "t = " << sv_physical_time->getValue() << "  TotalMechanicalEnergy = " << record_water_mechanical_energy.exec() << "\n";
    }
    execution_instance().synchronize();
    double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
    {
        const auto &last = fluid_observer_pressure.records().back();
        std::cout << "probe pressures at t = " << sv_physical_time->getValue() << ":";
        for (Real p : last) std::cout << " " << p;
        std::cout << "  (" << fluid_observer_pressure.records().size() << " records)\n";
    }
    std::cout << "Total wall time for computation: " << seconds << " seconds; "
              << double(water_block.TotalRealParticles()) * double(acoustic_steps) / seconds << " particle-steps/s\n";
    return 0;
}
