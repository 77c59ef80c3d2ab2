// examples/interpolation_restoring_2d.cpp — the reference's known-answer test of the first-order consistent interpolation,
// tests/unit_tests_src/shared/particle_dynamics/general_dynamics/unit_test_interpolation_ck/2d_interpolation.cpp:21-106
// (SYCL twin: tests/tests_sycl/unit_test_src/.../unit_test_interpolation_sycl/2d_interpolation.cpp), written against the
// C++ host layer in the statement order of its main(): a randomised lattice of fluid particles, one observer at a random
// point, ObservedQuantityRecording<Policy, Vecd, RestoringCorrection> of "Position" — the interpolated position must be the
// observer's own (linear reproduction). The reference runs Real = double and expects 1e-6; the device path is fp32: 1e-5.
// Build: g++ -O2 -std=c++17 -Iinclude examples/interpolation_restoring_2d.cpp -Lsphinxsys_b200 -lsphb200 -Wl,-rpath,$PWD/sphinxsys_b200
// Run: ./interpolation_restoring_2d [seed=1] [points=1]   (exit code 1 if any point misses the tolerance)
#include <iomanip>
#include <random>

#include "sphinxsys_ck/sphinxsys_ck.h"
using namespace SPH;

static Real rho0_f = 1.0, U_max = 1.0, c_f = 10.0 * U_max;          // 2d_interpolation.cpp:13-16
static Real width = 1.0, height = 0.5, particle_spacing = 0.01;     // :20-22
static Real boundary_width = particle_spacing * 4;

class WaterBlock : public ComplexShape
{
  public:
    explicit WaterBlock(const std::string &shape_name) : ComplexShape(shape_name)
    {
        Vecd scaled_container(0.5 * width, 0.5 * height);
        add<GeometricShapeBox>(Transform(scaled_container), scaled_container);
    }
};

int main(int ac, char *av[])
{
    const unsigned seed = ac > 1 ? (unsigned)std::atoi(av[1]) : 1u;
    const int points = ac > 2 ? std::atoi(av[2]) : 1;
    std::mt19937 gen(seed);
    auto rand_uniform = [&](Real lo, Real hi) { return std::uniform_real_distribution<Real>(lo, hi)(gen); };
    BoundingBoxd system_domain_bounds(Vecd(-boundary_width * 2, -boundary_width * 2), Vecd(width + boundary_width * 2, height + boundary_width * 2));
    SPHSystem sph_system(system_domain_bounds, particle_spacing, 2);
    //	Creating bodies with corresponding materials and particles.
    FluidBody water_block(sph_system, makeShared<WaterBlock>("WaterBody"));
    water_block.defineMatterMaterial<WeaklyCompressibleFluid>(rho0_f, c_f);
    water_block.generateParticles<BaseParticles, Lattice>();
    {
        // SimpleDynamics<relax_dynamics::RandomizeParticlePosition>::exec(0.5): every particle moved by up to a quarter spacing
        // per axis (relaxation pre-processing is host work, outside the hot path: done on the host view of the variable)
        auto *dv_pos = water_block.getBaseParticles().getVariableByName<Vecd>("Position");
        dv_pos->synchronizeWithDevice();
        Vecd *pos = dv_pos->Data();
        for (size_t i = 0; i != water_block.getBaseParticles().TotalRealParticles(); ++i)
        {
            pos[i].x += Real(0.5) * Real(0.5) * particle_spacing * rand_uniform(-1.0, 1.0);
            pos[i].y += Real(0.5) * Real(0.5) * particle_spacing * rand_uniform(-1.0, 1.0);
        }
        dv_pos->synchronizeToDevice();
        water_block.setPosVolDirty();
    }
    ObserverBody fluid_observer(sph_system, "FluidObserver");
    StdVec<Vecd> observation_location;
    for (int k = 0; k < points; ++k) observation_location.push_back(Vecd(rand_uniform(0.0, width), rand_uniform(0.0, height)));
    const StdVec<Vecd> reference_coordinate = observation_location;
    fluid_observer.generateParticles<ObserverParticles>(observation_location);
    //	Define body relation map.
    Contact<> fluid_observer_contact(fluid_observer, {&water_block});
    //	Define the numerical methods used in the simulation.
    UpdateCellLinkedList<MainExecutionPolicy, RealBody> water_cell_linked_list(water_block);
    UpdateRelation<MainExecutionPolicy, Contact<>> fluid_observer_contact_relation(fluid_observer_contact);
    ObservedQuantityRecording<MainExecutionPolicy, Vecd, RestoringCorrection> fluid_observer_position(fluid_observer_contact, "Position");
    //	Prepare the simulation with cell linked list and configuration.
    water_cell_linked_list.exec();
    fluid_observer_contact_relation.exec();
    fluid_observer_position.writeToFile(0);
    const Vecd *approximated_coordinate = fluid_observer_position.getObservedQuantity();
    Real worst = 0;
    for (int k = 0; k < points; ++k)
    {
        const Real ex = approximated_coordinate[k].x - reference_coordinate[k].x, ey = approximated_coordinate[k].y - reference_coordinate[k].y;
        worst = std::max(worst, Real(std::sqrt(ex * ex + ey * ey)));
    }
    std::cout << std::setprecision(9) << "Reference Coordinate: " << reference_coordinate[0].x << " " << reference_coordinate[0].y
              << " and Predicted Coordinate: " << approximated_coordinate[0].x << " " << approximated_coordinate[0].y << "\n"
              << "InterpolationError = " << worst << "\n";
    return worst < Real(1.0e-5) ? 0 : 1;
}
