// sphinxsys_ck/io_ck.h — the device -> host synchronisation hook of the output path.
//
// Reference: shared_ck/io_system/io_base_ck.h:41-58, io_base_ck.hpp:12-45 (BodyStatesRecordingToVtpCK::prepareToWrite:
// Position and every variable on the write list are brought to the host with prepareForOutput(policy) right before the
// file is written), io_system/io_base.h:95-111 (addToWrite<T>(body, name)), common/sphinxsys_variable.h:350-352
// (DiscreteVariable::prepareForOutput(ParallelDevicePolicy) = synchronizeWithDevice()).
// What is on the hot path is the HOOK: which arrays cross to the host, when, and in which particle order (the
// reference's numbering, through ReferenceID). The file writer behind it is deliberately small — one ASCII .vtp PolyData
// per body and call (points + the listed point data); the reference's binary/base64 VTP encoder is I/O, out of scope.
#ifndef SPHINXSYS_CK_IO_CK_H
#define SPHINXSYS_CK_IO_CK_H

#include <fstream>
#include <functional>
#include <iomanip>
#include <sstream>

#include "particles.h"

namespace SPH
{
template <class ExecutionPolicy> class BodyStatesRecordingToVtpCK
{
    struct Item
    {
        std::string name;
        int width;                          // 1 Real/int, 3 Vecd
        std::function<void()> prepare;      // prepareForOutput(ExecutionPolicy{})
        std::function<double(size_t, int)> value; // host value of particle i, component c (after prepare)
    };
    struct BodyEntry
    {
        SPHBody *body;
        std::vector<Item> items;
    };
    SPHSystem &sph_system_;
    std::vector<BodyEntry> bodies_;
    std::string output_folder_;
    bool state_recording_ = true;
    size_t files_written_ = 0, bytes_synchronized_ = 0;

    BodyEntry &entry(SPHBody &body)
    {
        for (auto &e : bodies_)
            if (e.body == &body) return e;
        throw SphError("BodyStatesRecording: the body '" + body.Name() + "' is not in the recording body list");
    }

  public:
    // all bodies of the system are recorded (io_base.h: BodyStatesRecording(SPHSystem &))
    explicit BodyStatesRecordingToVtpCK(SPHSystem &sph_system, const std::string &output_folder = "./output")
        : sph_system_(sph_system), output_folder_(output_folder)
    {
        execution::require_device_policy<ExecutionPolicy>();
        for (SPHBody *b : sph_system.bodies_) bodies_.push_back({b, {}});
    }
    void setStateRecording(bool on) { state_recording_ = on; }
    size_t filesWritten() const { return files_written_; }
    size_t bytesSynchronized() const { return bytes_synchronized_; }

    template <class T> BodyStatesRecordingToVtpCK &addToWrite(SPHBody &body, const std::string &name)
    {
        BaseParticles &p = body.getBaseParticles();
        DiscreteVariable<T> *v = p.template getVariableByName<T>(name);
        Item it;
        it.name = name;
        it.width = std::is_same<T, Vecd>::value ? 3 : 1;
        it.prepare = [v] { v->prepareForOutput(ExecutionPolicy{}); };
        it.value = [v](size_t i, int c) { return hostComponent(v->Data()[i], c); };
        entry(body).items.push_back(it);
        return *this;
    }

    // io_base_ck.hpp:12-24: positions and the write list of every body, device -> host, reference particle order
    void prepareToWrite()
    {
        for (auto &e : bodies_)
        {
            BaseParticles &p = e.body->getBaseParticles();
            p.template getVariableByName<Vecd>("Position")->prepareForOutput(ExecutionPolicy{});
            bytes_synchronized_ += p.hostSyncCount() * sizeof(Vecd);
            for (auto &it : e.items)
            {
                it.prepare();
                bytes_synchronized_ += p.hostSyncCount() * (it.width == 3 ? sizeof(Vecd) : sizeof(Real));
            }
        }
    }
    void writeToFile() { writeToFile(files_written_); }
    void writeToFile(size_t iteration_step)
    {
        if (!state_recording_) return;
        prepareToWrite();
        for (auto &e : bodies_) writeBody(e, iteration_step);
        ++files_written_;
    }

  private:
    static double hostComponent(const Real &v, int) { return v; }
    static double hostComponent(const int &v, int) { return v; }
    static double hostComponent(const UnsignedInt &v, int) { return v; }
    static double hostComponent(const Vecd &v, int c) { return c == 0 ? v.x : (c == 1 ? v.y : v.z); }

    void writeBody(BodyEntry &e, size_t iteration_step)
    {
        BaseParticles &p = e.body->getBaseParticles();
        const size_t n = p.hostSyncCount();
        std::ostringstream name;
        name << output_folder_ << "/" << e.body->Name() << "_" << std::setw(10) << std::setfill('0') << iteration_step << ".vtp";
        std::ofstream out(name.str());
        if (!out) throw SphError("BodyStatesRecording: cannot write " + name.str() + " (does the output folder exist?)");
        const Vecd *pos = p.template getVariableByName<Vecd>("Position")->Data();
        out << std::setprecision(9);
        out << "<?xml version=\"1.0\"?>\n<VTKFile type=\"PolyData\" version=\"0.1\" byte_order=\"LittleEndian\">\n<PolyData>\n"
            << "<Piece NumberOfPoints=\"" << n << "\" NumberOfVerts=\"0\" NumberOfLines=\"0\" NumberOfStrips=\"0\" NumberOfPolys=\"0\">\n"
            << "<Points>\n<DataArray type=\"Float32\" NumberOfComponents=\"3\" format=\"ascii\">\n";
        for (size_t i = 0; i < n; ++i) out << pos[i].x << " " << pos[i].y << " " << pos[i].z << "\n";
        out << "</DataArray>\n</Points>\n<PointData>\n";
        for (auto &it : e.items)
        {
            out << "<DataArray type=\"Float32\" Name=\"" << it.name << "\" NumberOfComponents=\"" << it.width << "\" format=\"ascii\">\n";
            for (size_t i = 0; i < n; ++i)
            {
                for (int c = 0; c < it.width; ++c) out << (c ? " " : "") << it.value(i, c);
                out << "\n";
            }
            out << "</DataArray>\n";
        }
        out << "</PointData>\n</Piece>\n</PolyData>\n</VTKFile>\n";
    }
};
} // namespace SPH
#endif
