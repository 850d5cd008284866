// parser.h -- string equations -> per-field implicit monomials + grouped explicit terms.
// Grammar and results are those of /root/reference/src/parser.cpp; the parser is the
// emitter of the fused per-equation plan: everything it produces is forwarded to the engine at prepareProblem.
#ifndef CUPSS_B200_PARSER_H
#define CUPSS_B200_PARSER_H

#include <map>
#include <string>
#include <vector>
#include "defines.h"

class evolver;

class parser {
   public:
    explicit parser(evolver *system);
    int createFromFile(const std::string &path);
    int add_equation(const std::string &equation);
    pres add_noise(const std::string &expression);
    int insert_parameter(const std::string &name, float value);
    int exists_parameter(const std::string &name);
    void writeParamsToFile(const std::string &path);
    float getParameter(const std::string &name);
    int changeParameter(const std::string &name, float value);
    int isParameterInString(const std::string &term, const std::string &parameter);
    int recalculateImplicits(const std::vector<std::string> &strings, std::vector<pres> &out, int dynamic);

   private:
    evolver *system;
    bool verbose = false;
    std::map<std::string, float> parameters;

    // factor classification
    int field_power(const std::string &factor);                       // 0 if not a field
    int q_power(const std::string &factor);                           // "q^n" -> n
    int tagged_power(const std::string &factor, const std::string &tag);   // "iqx", "iqx^n", "1/q", "1/q^n"
    bool looks_numeric(const std::string &text);
    float numeric_value(const std::string &factor);
    std::string field_of_factor(const std::string &factor);

    // term manipulation
    static std::string strip_spaces(const std::string &s);
    static void split_sides(const std::string &eq, std::string &lhs, std::string &rhs);
    static void split_sum(const std::string &s, std::vector<std::string> &terms);
    static void split_product(const std::string &s, std::vector<std::string> &factors);
    static bool distribute_once(const std::string &t, std::vector<std::string> &out);
    static void distribute_all(std::vector<std::string> &terms);
    static std::string normalise_division(const std::string &t);

    pres prefactor_of(const std::string &term);
    void fields_of(const std::string &term, std::vector<std::string> &fields);
    int count_fields(const std::string &term);
    std::string lhs_field(const std::string &term);
};

#endif
