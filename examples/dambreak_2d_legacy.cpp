// examples/dambreak_2d_legacy.cpp — the 2-D dam break written with the first-generation API names, in the statement
// order of the reference case file tests/2d_examples/test_2d_dambreak/Dambreak.cpp:60-220 (BASELINE.json config 1).
// Build: g++ -O2 -std=c++17 -Iinclude examples/dambreak_2d_legacy.cpp -Lsphinxsys_b200 -lsphb200 -Wl,-rpath,$PWD/sphinxsys_b200 -o dambreak_2d_legacy
#include <iomanip>

#include "sphinxsys_ck/legacy_dynamics.h"
using namespace SPH;

static Real DL = 5.366, DH = 5.366, LL = 2.0, LH = 1.0; // Dambreak.cpp:13-16
static Real particle_spacing_ref = 0.025;
static Real rho0_f = 1.0, gravity_g = 1.0;

class WaterBlock : public ComplexShape
{
  public:
    explicit WaterBlock(const std::string &shape_name) : ComplexShape(shape_name)
    {
        Vecd water_block_halfsize(0.5 * LL, 0.5 * LH);
        add<GeometricShapeBox>(Transform(water_block_halfsize), water_block_halfsize);
    }
};
class WallBoundary : public ComplexShape
{
  public:
    explicit WallBoundary(const std::string &shape_name) : ComplexShape(shape_name)
    {
        Real BW = particle_spacing_ref * 4;
        Vecd inner_wall_halfsize(0.5 * DL, 0.5 * DH);
        Vecd outer_wall_halfsize(0.5 * DL + BW, 0.5 * DH + BW);
        add<GeometricShapeBox>(Transform(inner_wall_halfsize), outer_wall_halfsize);
        subtract<GeometricShapeBox>(Transform(inner_wall_halfsize), inner_wall_halfsize);
    }
};

int main(int ac, char *av[])
{
    if (ac > 1) particle_spacing_ref = Real(std::atof(av[1]));
    Real end_time = ac > 2 ? Real(std::atof(av[2])) : Real(1.0);
    Real BW = particle_spacing_ref * 4;
    Real U_ref = 2.0 * std::sqrt(gravity_g * LH), c_f = 10.0 * U_ref;
    BoundingBoxd system_domain_bounds(Vecd(-BW, -BW), Vecd(DL + BW, DH + BW));
    SPHSystem sph_system(system_domain_bounds, particle_spacing_ref, 2);
    //	Creating bodies with corresponding materials and particles.
    FluidBody water_block(sph_system, makeShared<WaterBlock>("WaterBody"));
    water_block.defineMatterMaterial<WeaklyCompressibleFluid>(rho0_f, c_f);
    water_block.generateParticles<BaseParticles, Lattice>();
    SolidBody wall_boundary(sph_system, makeShared<WallBoundary>("WallBoundary"));
    wall_boundary.defineMatterMaterial<Solid>();
    wall_boundary.generateParticles<BaseParticles, Lattice>();
    //	Define body relation map.
    InnerRelation water_block_inner(water_block);
    ContactRelation water_wall_contact(water_block, {&wall_boundary});
    ComplexRelation water_wall_complex(water_block_inner, water_wall_contact);
    //	Define the numerical methods used in the simulation.
    Gravity gravity(Vecd(0.0, -gravity_g));
    SimpleDynamics<GravityForce<Gravity>> constant_gravity(water_block, gravity);
    Dynamics1Level<fluid_dynamics::Integration1stHalfWithWallRiemann> fluid_pressure_relaxation(water_block_inner, water_wall_contact);
    Dynamics1Level<fluid_dynamics::Integration2ndHalfWithWallRiemann> fluid_density_relaxation(water_block_inner, water_wall_contact);
    InteractionWithUpdate<fluid_dynamics::DensitySummationComplexFreeSurface> fluid_density_by_summation(water_block_inner, water_wall_contact);
    ReduceDynamics<fluid_dynamics::AdvectionViscousTimeStep> fluid_advection_time_step(water_block, U_ref);
    ReduceDynamics<fluid_dynamics::AcousticTimeStep> fluid_acoustic_time_step(water_block);
    ReduceDynamics<TotalMechanicalEnergy> write_water_mechanical_energy(water_block, gravity);
    ParticleSorting particle_sorting(water_block);
    //	Prepare the simulation with cell linked list, configuration and case specified initial condition.
    wall_boundary.computeNormalFromBodyShape();
    constant_gravity.exec();
    updateCellLinkedList(water_block);   // sph_system.initializeSystemCellLinkedLists()
    updateCellLinkedList(wall_boundary);
    water_wall_complex.updateConfiguration(); // sph_system.initializeSystemConfigurations()
    size_t number_of_iterations = 0;
    int screen_output_interval = 100;
    Real physical_time = 0, output_interval = end_time / 10.0;
    std::cout << "N_fluid = " << water_block.TotalRealParticles() << "  E0 = " << std::setprecision(9) << write_water_mechanical_energy.exec() << "\n";
    //	Main loop starts here.
    while (physical_time < end_time)
    {
        Real integration_time = 0.0;
        while (integration_time < output_interval)
        {
            Real advection_dt = fluid_advection_time_step.exec();
            fluid_density_by_summation.exec();
            Real relaxation_time = 0.0, acoustic_dt = 0.0;
            while (relaxation_time < advection_dt)
            {
                acoustic_dt = fluid_acoustic_time_step.exec();
                fluid_pressure_relaxation.exec(acoustic_dt);
                fluid_density_relaxation.exec(acoustic_dt);
                relaxation_time += acoustic_dt;
                integration_time += acoustic_dt;
                physical_time += acoustic_dt;
            }
            if (number_of_iterations % screen_output_interval == 0)
                std::cout << std::fixed << std::setprecision(9) << "N=" << number_of_iterations << "	Time = " << physical_time
                          << "	advection_dt = " << advection_dt << "	acoustic_dt = " << acoustic_dt << "\n";
            number_of_iterations++;
            if (number_of_iterations % 100 == 0 && number_of_iterations != 1) particle_sorting.exec();
            updateCellLinkedList(water_block); // water_block.updateCellLinkedList()
            water_wall_complex.updateConfiguration();
        }
        std::cout << "t = " << physical_time << "  TotalMechanicalEnergy = " << write_water_mechanical_energy.exec() << "\n";
    }
    return 0;
}
