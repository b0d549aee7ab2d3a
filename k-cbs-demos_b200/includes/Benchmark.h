// Benchmark.h -- mirror of the reference's includes/Benchmark.h (:93-230): runs one K-CBS or PP solve, times it with
// std::chrono and appends one row to <name>.csv (columns in std::map key order, header only when the file is new).
#pragma once
#include <chrono>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <string>
#include <vector>

#include <ompl/multirobot/base/Planner.h>
#include <ompl/multirobot/base/ProblemDefinition.h>
#include <ompl/multirobot/control/SpaceInformation.h>
#include <ompl/multirobot/control/planners/kcbs/KCBS.h>
#include <ompl/multirobot/control/planners/pp/PP.h>

namespace omrb = ompl::multirobot::base;
namespace omrc = ompl::multirobot::control;

class Benchmark
{
public:
    Benchmark() : si_(nullptr), pdef_(nullptr), solveTime_(300), numRuns_(100), filename_("Results") {}
    ~Benchmark()
    {
        results_.clear();
        si_.reset();
        pdef_.reset();
        for (auto &p : planners_) {
            std::cout << p.use_count() << std::endl;
            p.reset();
        }
    }

    void setSpaceInformation(omrc::SpaceInformationPtr &si) { si_ = si; }
    void setPoroblemDefinition(omrb::ProblemDefinitionPtr &pdef) { pdef_ = pdef; }  // (sic) reference spelling, Benchmark.h:119
    void addPlanner(omrb::PlannerPtr &p) { planners_.push_back(p); }
    void setSolveTime(const double t) { solveTime_ = t; }
    void setKCBSMergeBound(int value) { merge_bound_ = value; }
    void setNumberOfRuns(const unsigned int n) { numRuns_ = n; }
    void setFileName(const std::string name) { filename_ = name; }

    void runKCBS()
    {
        omrc::KCBS p(si_);
        p.setProblemDefinition(pdef_);
        p.setLowLevelSolveTime(5.);
        p.setMergeBound(merge_bound_);
        double seconds = 0;
        const ompl::base::PlannerStatus solved = timedSolve(p, seconds);
        results_.insert({"K-CBS Success (Bool)", std::to_string(solved == ompl::base::PlannerStatus::EXACT_SOLUTION)});
        results_.insert({"K-CBS Computation Times (s)", std::to_string(seconds)});
        results_.insert({"K-CBS Root Node Solve Time (s)", std::to_string(p.getRootSolveTime())});
        results_.insert({"K-CBS Number of Expanded Nodes", std::to_string(p.getNumberOfNodesExpanded())});
        results_.insert({"K-CBS Number Approximate Solutions", std::to_string(p.getNumberOfApproximateSolutions())});
        pdef_->clearSolutionPaths();
        p.clear();
        writeCSV();
    }

    void runPP()
    {
        omrc::PP p(si_);
        p.setProblemDefinition(pdef_);
        double seconds = 0;
        const ompl::base::PlannerStatus solved = timedSolve(p, seconds);
        results_.insert({"PP Success (Bool)", std::to_string(solved == ompl::base::PlannerStatus::EXACT_SOLUTION)});
        results_.insert({"PP Computation Times (s)", std::to_string(seconds)});
        pdef_->clearSolutionPaths();
        p.clear();
        writeCSV();
    }

    void writeCSV()
    {
        const std::string name = filename_ + ".csv";
        bool exist;
        {
            std::ifstream infile(name);
            exist = infile.good();
        }
        if (!exist) {
            std::ofstream heads(name);
            for (const auto &kv : results_) heads << kv.first << ",";
            heads << std::endl;
        }
        std::ofstream stats(name, std::ios::app);
        for (const auto &kv : results_) stats << kv.second << ",";
        stats << std::endl;
    }

private:
    // wall-clock of one solve, in whole milliseconds like the reference (Benchmark.h:157-161)
    template <class P>
    ompl::base::PlannerStatus timedSolve(P &planner, double &seconds)
    {
        const auto start = std::chrono::high_resolution_clock::now();
        const ompl::base::PlannerStatus solved = planner.template as<omrb::Planner>()->solve(solveTime_);
        const auto end = std::chrono::high_resolution_clock::now();
        seconds = std::chrono::duration_cast<std::chrono::milliseconds>(end - start).count() * 0.001;
        return solved;
    }

    omrc::SpaceInformationPtr si_;
    omrb::ProblemDefinitionPtr pdef_;
    std::vector<omrb::PlannerPtr> planners_;
    double solveTime_;
    unsigned int numRuns_;
    int merge_bound_ = std::numeric_limits<int>::max();
    std::string filename_;
    std::map<std::string, std::string> results_;
};
