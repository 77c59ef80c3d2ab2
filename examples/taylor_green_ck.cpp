// examples/taylor_green_ck.cpp — the periodic Taylor-Green vortex (BASELINE config 4) on the C++ host layer.
// Case set-up and loop order follow tests/2d_examples/test_2d_taylor_green/taylor_green.cpp:62-200 in its CK spelling
// (include/sphinxsys_ck/taylor_green_case.h), in 2-D or 3-D. With `ring` the x axis is periodic through the slab
// exchange of a decomposed run (here: a ring of ONE slab on a communicator without NCCL) instead of
// through image particles; an N-GPU launcher hands every rank its own particles (sphinxsys_b200/host.py does).
// Build: g++ -O2 -std=c++17 -Iinclude examples/taylor_green_ck.cpp -Lsphinxsys_b200 -lsphb200 -Wl,-rpath,$PWD/sphinxsys_b200 -o taylor_green_ck
// Run:   ./taylor_green_ck [n_side=64] [dim=3] [end_time=0.2] [ring=0] [Re=0 (inviscid); > 0: viscous + transport velocity]
#include <chrono>
#include <iomanip>

#include "sphinxsys_ck/taylor_green_case.h"
using namespace SPH;

int main(int ac, char *av[])
{
    const int n_side = ac > 1 ? std::atoi(av[1]) : 64;
    const int dim = ac > 2 ? std::atoi(av[2]) : 3;
    const double end_time = ac > 3 ? std::atof(av[3]) : 0.2;
    const bool ring = ac > 4 && std::atoi(av[4]) != 0;
    const double Re = ac > 5 ? std::atof(av[5]) : 0.0;
    try
    {
        TaylorGreenParameters q;
        q.dim = dim;
        q.dp = 1.0 / n_side;        // global_resolution = DL / particle number per side (taylor_green.cpp:14-16)
        q.sort_interval = 100;
        if (Re > 0)
        {
            q.mu_f = q.rho0_f * q.U_f * q.L / Re; // taylor_green.cpp:19-22
            q.transport_velocity = true;
        }
        std::unique_ptr<TaylorGreenCK> sim;
        if (ring)
        {
            // a ring takes its particles from the caller: the lattice of the box and the analytic initial condition
            ExecutionInstance &ex = execution_instance();
            ex.check(sphb200_comm_create_self(ex.ctx()), "sphb200_comm_create_self");
            ex.check(sphb200_comm_set_ring(ex.ctx(), 1), "sphb200_comm_set_ring");
            q.ring = true;
            std::vector<Vecd> pos, vel;
            const int nz = dim == 3 ? n_side : 1;
            for (int i = 0; i < n_side; ++i)
                for (int j = 0; j < n_side; ++j)
                    for (int k = 0; k < nz; ++k)
                    {
                        Vecd x(Real((i + 0.5) * q.dp), Real((j + 0.5) * q.dp), dim == 3 ? Real((k + 0.5) * q.dp) : Real(0));
                        pos.push_back(x);
                        vel.push_back(TaylorGreenCK::initialVelocity(x, dim, Real(q.U_f)));
                    }
            sim.reset(new TaylorGreenCK(q, &pos, &vel));
        }
        else
            sim.reset(new TaylorGreenCK(q));
        sim->initialize();
        const size_t n = sim->water_block.getBaseParticles().activeEnd() - sim->water_block.getBaseParticles().activeBegin();
        std::cout << "Taylor-Green " << dim << "-D, " << n << " particles, " << (ring ? "ring of one slab" : "periodic images")
                  << (Re > 0 ? ", viscous + transport velocity" : ", inviscid") << std::endl;
        const Real e0 = sim->record_total_kinetic_energy->exec();
        auto t0 = std::chrono::steady_clock::now();
        while (sim->physical_time < end_time)
        {
            sim->stepOuter();
            if (sim->number_of_iterations % 50 == 0)
                std::cout << std::fixed << std::setprecision(6) << "N=" << sim->number_of_iterations << "  Time = " << sim->physical_time
                          << "  advection_dt = " << sim->last_advection_dt << "  acoustic_dt = " << sim->last_acoustic_dt << std::endl;
        }
        execution_instance().synchronize();
        const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const Real e1 = sim->record_total_kinetic_energy->exec();
        std::cout << std::setprecision(6) << "kinetic energy " << e0 << " -> " << e1 << " at t = " << sim->physical_time << "; "
                  << sim->number_of_iterations << " advection / " << sim->acoustic_steps << " acoustic steps in " << wall << " s = "
                  << std::scientific << std::setprecision(3) << double(n) * double(sim->acoustic_steps) / wall << " particle-steps/s" << std::endl;
        sim.reset();
        if (ring) sphb200_comm_destroy(execution_instance().ctx());
    }
    catch (const std::exception &e)
    {
        std::cout << "\n Error: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
